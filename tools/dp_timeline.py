"""Where a data-parallel step spends its time: CUDA events on the main stream after every backward stage, on an observer
stream after every gradient all-reduce, and around the optimizer.  Run under torchrun (rank 0 prints), e.g.
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29560 tools/dp_timeline.py
Prints one JSON line: per-stage end times (ms after the start of backward), per-slice all-reduce end times, the wait the
main stream sees after the last stage, the optimizer time and the step time."""
import argparse, json, os, sys
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_ae_plus_plus_b200 import dp                                    # noqa: E402
from vit_ae_plus_plus_b200.model import model_factory                    # noqa: E402
from vit_ae_plus_plus_b200.utils import misc                             # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--batch", type=int, default=4)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dry = bool(os.environ.get("DP_TIMELINE_DRY"))
    if not dry:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    import bench
    a.workload, a.mask_ratio, a.edge_map_weight = "vit_base_128", 0.75, 0.0
    margs = bench.model_args(bench.WORKLOADS[a.workload], a)
    torch.manual_seed(42 + rank)
    model = model_factory.get_models("autoenc", margs)
    if os.environ.get("DP_TIMELINE_DRY"):
        print("constructed", type(model).__name__)
        return
    model = model.to(dev)
    model.train(True)
    model.pred_dtype = torch.bfloat16
    opt = torch.optim.AdamW(misc.add_weight_decay(model, 0.05), lr=1e-4, betas=(0.9, 0.95))
    scaler = misc.NativeScalerWithGradNormCount()
    pool = [torch.randn(a.batch, 4, 128, 128, 128, device=dev) for _ in range(2)]
    eng = model.engine()
    obs = torch.cuda.Stream()
    rec = {"stage": [], "ar": [], "marks": {}}

    def ev(stream=None):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream or torch.cuda.current_stream())
        return e

    orig_run = eng._run
    def run(pl, key, fn):
        r = orig_run(pl, key, fn)
        if key[0] == "bwd":
            if key[1] == 0:
                pass
            rec["stage"].append(ev())
        return r
    eng._run = run

    orig_launch = dp.GradReducer.launch
    def launch(self, t):
        n0 = len(self.pending)
        orig_launch(self, t)
        if len(self.pending) > n0:
            work = self.pending[-1][0]
            with torch.cuda.stream(obs):
                if work is not None:
                    work.wait()
                else:
                    obs.wait_stream(self.stream)
                rec["ar"].append((ev(obs), t.numel()))
    dp.GradReducer.launch = launch
    orig_wait = dp.GradReducer.wait
    def wait(self):
        orig_wait(self)
        rec["marks"]["after_wait"] = ev()
    dp.GradReducer.wait = wait

    # the sharded step's reducer (dp.ShardedStep): same marks, plus the step's own phases
    orig_slaunch = dp.ShardedStep.launch
    def slaunch(self, t, overlap=True, final=False):
        orig_slaunch(self, t, overlap, final)
        with torch.cuda.stream(obs):
            obs.wait_stream(self.comm if overlap else torch.cuda.current_stream())
            rec["ar"].append((ev(obs), t.numel()))
    dp.ShardedStep.launch = slaunch
    orig_swait = dp.ShardedStep.wait
    def swait(self):
        orig_swait(self)
        rec["marks"]["after_wait"] = ev()
    dp.ShardedStep.wait = swait

    orig_bwd = eng.backward
    def backward(*args, **kw):
        rec["marks"]["bwd0"] = ev()
        return orig_bwd(*args, **kw)
    eng.backward = backward

    def step(x):
        rec["stage"], rec["ar"], rec["marks"] = [], [], {}
        rec["marks"]["t0"] = ev()
        losses, _p, _m = model(x, mask_ratio=0.75, edge_map_weight=0.0)
        scaler(losses[0], opt, parameters=model.parameters(), update_grad=True)
        rec["marks"]["opt_end"] = ev()
        opt.zero_grad()
        rec["marks"]["t1"] = ev()

    for _ in range(4):
        for x in pool:
            step(x)
    acc = None
    for i in range(a.steps):
        step(pool[i % 2])
        torch.cuda.synchronize()
        m = rec["marks"]
        b0 = m["bwd0"]
        row = {"fwd": m["t0"].elapsed_time(b0), "stage_end": [b0.elapsed_time(e) for e in rec["stage"]],
               "ar_end": [b0.elapsed_time(e) for e, _ in rec["ar"]], "ar_mb": [n * 4 / 1e6 for _, n in rec["ar"]],
               "after_wait": b0.elapsed_time(m["after_wait"]) if "after_wait" in m else None,
               "opt_end": b0.elapsed_time(m["opt_end"]), "step": m["t0"].elapsed_time(m["t1"])}
        if acc is None:
            acc = {k: ([0.0] * len(v) if isinstance(v, list) else 0.0) for k, v in row.items() if v is not None}
        for k, v in row.items():
            if v is None:
                continue
            if isinstance(v, list):
                acc[k] = [p + q for p, q in zip(acc[k], v)]
            else:
                acc[k] += v
    out = {k: ([round(x / a.steps, 3) for x in v] if isinstance(v, list) else round(v / a.steps, 3)) for k, v in acc.items()}
    out["world"] = world
    out["sharded"] = bool(eng.flat.sharded is not None and eng.flat.sharded.steps > 0)
    out["env"] = {k: os.environ[k] for k in os.environ if k.startswith(("NCCL_", "VITAE_"))}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
