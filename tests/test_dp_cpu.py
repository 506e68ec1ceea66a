"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2 (the GPU box runs the same code over NCCL):
flat-buffer broadcast, bucketed mean all-reduce, the loop's deferred scalar reduction and the meters' synchronisation."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vit_ae_plus_plus_b200 import dp
        from vit_ae_plus_plus_b200.engine import backward_param_order
        from vit_ae_plus_plus_b200.utils import misc
        from vit_ae_plus_plus_b200.utils.train_one_epoch import _DeferredScalars
        out = {}
        # parameters: every rank seeds differently (k_fold_..._brats.py:87); rank 0 wins
        torch.manual_seed(42 + rank)
        p = torch.randn(1000)
        dp.broadcast_flat(p)
        out["p_sum"] = p.sum().item()
        # gradients: bucketed mean all-reduce over slices that end on tensor boundaries
        sizes = [(0, 300), (300, 10), (310, 500), (810, 190)]
        slices = dp.bucket_slices(sizes, 1000, 400)
        out["slices"] = slices
        g = torch.full((1000,), float(rank + 1))
        g[5] = 10.0 * (rank + 1)
        dp.allreduce_mean_bucketed_(g, slices)
        out["g0"], out["g5"] = g[0].item(), g[5].item()
        # asynchronous per-stage reducer (engine.backward(sync_grads=True)): slices in flight together, then one wait
        g2 = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        red = dp.GradReducer()
        for a, b in [(0, 128), (128, 640), (640, 1000)]:
            red.launch(g2[a:b])
        red.wait()
        out["g2_ok"] = bool(torch.equal(g2, torch.arange(1000, dtype=torch.float32) * 1.5)) and red.pending == []
        out["all_reduce_mean"] = misc.all_reduce_mean(float(rank))
        sv = misc.SmoothedValue()
        sv.update(1.0 + rank, n=2)
        sv.synchronize_between_processes()
        out["sv"] = (sv.count, sv.total)

        class W:
            def __init__(self):
                self.rows = []

            def add_scalar(self, tag, v, x):
                self.rows.append((tag, v, x))
        w = W()
        ml = misc.MetricLogger()
        d = _DeferredScalars(["edge_map_loss", "reconstruction_loss", "perceptual_loss", "contr_loss", "loss"], ml, w,
                             lambda v: (_ for _ in ()).throw(RuntimeError("nonfinite")))
        d.add([torch.tensor(0.), torch.tensor(1.0 + rank), torch.tensor(0.), torch.tensor(0.), torch.tensor(1.0 + rank)], 1e-3, 7, True)
        d.flush()
        out["local_loss"] = ml.meters["loss"].global_avg
        out["logged_loss"] = [r for r in w.rows if r[0] == "train_loss"][0][1]
        out["order_ok"] = backward_param_order(2, 1)[0] == "decoder_pred.weight" and backward_param_order(2, 1)[-1] == "patch_embed.proj.bias"
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert res[0]["p_sum"] == res[1]["p_sum"]                       # replicas identical after the broadcast
    assert res[0]["slices"] == [(0, 810), (810, 1000)] or res[0]["slices"][0][0] == 0
    assert res[0]["slices"][-1][1] == 1000
    for r in (0, 1):
        assert res[r]["g0"] == 1.5 and res[r]["g5"] == 15.0         # mean over ranks
        assert res[r]["g2_ok"]
        assert res[r]["all_reduce_mean"] == 0.5
        assert res[r]["sv"] == (4, 6.0)
        assert res[r]["local_loss"] == 1.0 + r                      # meters stay rank-local (reference misc.py:95)
        assert res[r]["logged_loss"] == 1.5                         # TensorBoard gets the rank mean (train_one_epoch.py:83-88)
        assert res[r]["order_ok"]


def test_bucket_slices_cover_everything():
    from vit_ae_plus_plus_b200 import dp
    offs, o = [], 0
    for k in [5, 64, 1000, 3, 3, 700, 64, 64]:
        offs.append((o, k))
        o += (k + 63) // 64 * 64
    sl = dp.bucket_slices(offs, o, 512)
    assert sl[0][0] == 0 and sl[-1][1] == o
    assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
    assert dp.bucket_slices(offs, o, 1 << 30) == [(0, o)]
