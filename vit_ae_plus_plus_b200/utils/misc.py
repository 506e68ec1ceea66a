"""The pieces of the reference's ``utils/misc.py`` that the MAE pre-training loop touches, restated for a loop that does
not synchronise the device every step: windowed meters (misc.py:24-100), the iteration logger (:102-167), the
GradScaler wrapper the k-fold scripts construct (:251-277), the gradient norm (:280-292), the scalar all-reduce
(:332-340) and the timm-0.5.4 weight-decay grouping used at the optimizer call site
(k_fold_cross_valid_combined_brats.py:168), plus the launch / checkpoint glue the k-fold scripts call on this module
(``init_distributed_mode`` :216-248 at brats.py:78, ``load_model`` :313-329 at :173, ``save_model`` :295-310 at :198,
``setup_for_distributed`` :170-184, ``save_on_master`` :211-213), so that an overlay which shadows ``utils.misc`` with
this module (overlay/utils/__init__.py) leaves no name of the reference's module unresolved."""
from __future__ import annotations

import builtins
import datetime
import os
import time
from collections import defaultdict, deque
from pathlib import Path

import torch
import torch.distributed as dist


def is_dist_avail_and_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_world_size() -> int:
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank() -> int:
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process() -> bool:
    return get_rank() == 0


def setup_for_distributed(is_master: bool) -> None:
    """Silences ``print`` on non-master ranks (reference misc.py:170-184): ``print(..., force=True)`` always prints, and
    so does every rank of a job with more than 8 ranks; printed lines carry a wall-clock stamp."""
    plain = getattr(builtins.print, "_vitae_plain", builtins.print)     # re-patching must not stack time stamps

    def print(*args, **kwargs):
        force = kwargs.pop("force", False) or get_world_size() > 8
        if is_master or force:
            plain("[{}] ".format(datetime.datetime.now().time()), end="")
            plain(*args, **kwargs)

    print._vitae_plain = plain
    builtins.print = print


def save_on_master(*args, **kwargs) -> None:
    if is_main_process():
        torch.save(*args, **kwargs)


def init_distributed_mode(args) -> None:
    """Reads the launcher's environment into ``args`` and brings up the NCCL process group (reference misc.py:216-248).
    Launch styles, in the reference's order: OpenMPI (``args.dist_on_itp``), torchrun (RANK / WORLD_SIZE / LOCAL_RANK),
    SLURM (SLURM_PROCID); none of them -> ``args.distributed = False`` and printing stays on.  One process drives one
    GPU; this package's models broadcast their parameters and average their gradients themselves once the group exists
    (vit_ae_plus_plus_b200/dp.py), so no DistributedDataParallel wrapper follows."""
    env = os.environ
    if getattr(args, "dist_on_itp", False):
        args.rank, args.world_size = int(env["OMPI_COMM_WORLD_RANK"]), int(env["OMPI_COMM_WORLD_SIZE"])
        args.gpu = int(env["OMPI_COMM_WORLD_LOCAL_RANK"])
        args.dist_url = "tcp://%s:%s" % (env["MASTER_ADDR"], env["MASTER_PORT"])
        env["LOCAL_RANK"], env["RANK"], env["WORLD_SIZE"] = str(args.gpu), str(args.rank), str(args.world_size)
    elif "RANK" in env and "WORLD_SIZE" in env:
        args.rank, args.world_size, args.gpu = int(env["RANK"]), int(env["WORLD_SIZE"]), int(env["LOCAL_RANK"])
    elif "SLURM_PROCID" in env:
        args.rank = int(env["SLURM_PROCID"])
        args.gpu = args.rank % torch.cuda.device_count()
    else:
        print("Not using distributed mode")
        setup_for_distributed(is_master=True)
        args.distributed = False
        return
    args.distributed = True
    if not hasattr(args, "dist_url"):
        args.dist_url = "env://"
    if not hasattr(args, "world_size"):
        args.world_size = int(env.get("WORLD_SIZE", env.get("SLURM_NTASKS", "1")))
    # the reference always asks for NCCL; VITAE_DIST_BACKEND=gloo lets the CPU tests exercise the same code
    args.dist_backend = env.get("VITAE_DIST_BACKEND", "nccl")
    if args.dist_backend == "nccl":
        torch.cuda.set_device(args.gpu)
    print("| distributed init (rank {}): {}, gpu {}".format(args.rank, args.dist_url, args.gpu), flush=True)
    dist.init_process_group(backend=args.dist_backend, init_method=args.dist_url, world_size=args.world_size,
                            rank=args.rank)
    dist.barrier()
    setup_for_distributed(args.rank == 0)


class SmoothedValue:
    """Window median / mean plus the global average of a scalar series (reference misc.py:24-84)."""

    def __init__(self, window_size=20, fmt=None):
        self.fmt = fmt or "{median:.4f} ({global_avg:.4f})"
        self.deque = deque(maxlen=window_size)
        self.total = 0.0
        self.count = 0

    def update(self, value, n=1):
        self.deque.append(value)
        self.count += n
        self.total += value * n

    def synchronize_between_processes(self):
        """Sums count / total over ranks (the window is left rank-local, like the reference)."""
        if not is_dist_avail_and_initialized():
            return
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([self.count, self.total], dtype=torch.float64, device=dev)
        dist.barrier()
        dist.all_reduce(t)
        self.count, self.total = int(t[0].item()), t[1].item()

    @property
    def median(self):
        return torch.tensor(list(self.deque)).median().item()

    @property
    def avg(self):
        return torch.tensor(list(self.deque), dtype=torch.float32).mean().item()

    @property
    def global_avg(self):
        return self.total / self.count

    @property
    def max(self):
        return max(self.deque)

    @property
    def value(self):
        return self.deque[-1]

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg, max=self.max,
                               value=self.value)


class MetricLogger:
    def __init__(self, delimiter="\t"):
        self.meters = defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def update(self, **kwargs):
        for name, v in kwargs.items():
            if v is None:
                continue
            if isinstance(v, torch.Tensor):
                v = v.item()
            assert isinstance(v, (float, int))
            self.meters[name].update(v)

    def __getattr__(self, attr):
        meters = self.__dict__.get("meters", {})
        if attr in meters:
            return meters[attr]
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{attr}'")

    def __str__(self):
        return self.delimiter.join(f"{name}: {meter}" for name, meter in self.meters.items())

    def synchronize_between_processes(self):
        for meter in self.meters.values():
            meter.synchronize_between_processes()

    def add_meter(self, name, meter):
        self.meters[name] = meter

    def log_every(self, iterable, print_freq, header=None, before_print=None):
        """Yields from ``iterable``; prints meters / eta / iteration time every ``print_freq`` items.  ``before_print`` is
        called just before a line is printed (our loop uses it to flush its deferred device scalars)."""
        header = header or ""
        n = len(iterable)
        iter_time, data_time = SmoothedValue(fmt="{avg:.4f}"), SmoothedValue(fmt="{avg:.4f}")
        width = len(str(n))
        start = end = time.time()
        for i, obj in enumerate(iterable):
            data_time.update(time.time() - end)
            yield obj
            iter_time.update(time.time() - end)
            if i % print_freq == 0 or i == n - 1:
                if before_print is not None:
                    before_print()
                eta = str(datetime.timedelta(seconds=int(iter_time.global_avg * (n - i))))
                parts = [header, f"[{i:{width}d}/{n}]", f"eta: {eta}", str(self), f"time: {iter_time}", f"data: {data_time}"]
                if torch.cuda.is_available():
                    parts.append(f"max mem: {torch.cuda.max_memory_allocated() / 2 ** 20:.0f}")
                print(self.delimiter.join(parts))
            end = time.time()
        total = time.time() - start
        print(f"{header} Total time: {datetime.timedelta(seconds=int(total))} ({total / max(n, 1):.4f} s / it)")


def normalize_volumes(raw: torch.Tensor, mode: str = "z_score_channel", out: torch.Tensor = None) -> torch.Tensor:
    """The reference's ``Dataset._normalize_data`` on the device: ``raw`` [B, C, V, V, V] CUDA tensor in its
    storage dtype (uint16 / int16 / uint8 / float16 / bfloat16 / float32) -> fp32 of the same shape.  ``mode``:
    "z_score_channel" (dataset/egd_dataset/egd.py:45-47: per channel, unbiased variance), "z_score_sample"
    (dataset/brats_dataset/brats.py:27-29: over the whole sample), "min_max" (:30-32 / egd.py:48-50: to [-1, 1])."""
    from .. import ops
    raw = raw.contiguous()
    if out is None:
        out = torch.empty(raw.shape, dtype=torch.float32, device=raw.device)
    ws = torch.empty(ops.ingest_workspace_bytes(raw.shape[0], raw.shape[1]), dtype=torch.uint8, device=raw.device)
    ops.ingest_normalize(raw, out, mode, ws)
    return out


class DevicePrefetcher:
    """Wraps a batch iterable (the k-fold scripts' DataLoader, k_fold_cross_valid_combined_brats.py:131-148): batch k+1 is
    copied host -> device on a dedicated copy stream while step k computes, into ``depth`` rotating device buffers per
    tensor slot (stable addresses: the step's CUDA graphs read the volume in place).  The reference loop's
    ``sample.to(device, non_blocking=True)`` (utils/train_one_epoch.py:47-48) then finds the tensors already resident.
    A 4 x 4 x 128^3 fp32 batch is 134 MB = ~2.6 ms of PCIe time per step, more than half a B200 training step; host
    tensors should be pinned (DataLoader(pin_memory=True), the scripts' default) or the copy cannot overlap.

    ``normalize`` (None | "z_score_channel" | "z_score_sample" | "min_max"): the loader yields RAW volumes in their storage
    dtype (uint16 / float16 ...: half the PCIe bytes of fp32, or fewer) and every 5-D tensor of a batch is normalised on the
    device right after its copy, on the copy stream (``normalize_volumes``: the reference does this on the host inside
    ``Dataset.__getitem__``); the consumer sees fp32 volumes as before."""

    def __init__(self, loader, device, depth: int = 2, normalize: str = None):
        self.loader, self.device, self.depth = loader, torch.device(device), max(2, int(depth))
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.bufs = [dict() for _ in range(self.depth)]
        self.normalize = normalize
        self.h2d_bytes = 0          # bytes copied host -> device so far (bench.py reports them)

    def __len__(self):
        return len(self.loader)

    def _issue(self, slot: int, batch, free_event):
        from .. import ops
        items = list(batch) if isinstance(batch, (tuple, list)) else [batch]
        with torch.cuda.stream(self.stream):
            if free_event is not None:
                self.stream.wait_event(free_event)      # the consumer has finished with this slot's previous batch
            out = []
            for j, t in enumerate(items):
                if not torch.is_tensor(t):
                    out.append(t)
                    continue
                buf = self.bufs[slot].get(j)
                if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                    buf = self.bufs[slot][j] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                buf.copy_(t, non_blocking=True)
                self.h2d_bytes += t.numel() * t.element_size()
                if self.normalize is not None and t.dim() == 5:
                    key = ("norm", j)
                    nb = self.bufs[slot].get(key)
                    if nb is None or nb[0].shape != t.shape:
                        nb = self.bufs[slot][key] = (
                            torch.empty(t.shape, dtype=torch.float32, device=self.device),
                            torch.empty(ops.ingest_workspace_bytes(t.shape[0], t.shape[1]), dtype=torch.uint8, device=self.device))
                    ops.ingest_normalize(buf, nb[0], self.normalize, nb[1])
                    buf = nb[0]
                out.append(buf)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return (tuple(out) if isinstance(batch, (tuple, list)) else out[0]), ready

    def __iter__(self):
        if self.stream is None:
            yield from self.loader
            return
        it = iter(self.loader)
        free = [None] * self.depth
        try:
            pending = self._issue(0, next(it), None)
        except StopIteration:
            return
        k = 0
        while pending is not None:
            cur, ready = pending
            try:
                nxt = next(it)
                pending = self._issue((k + 1) % self.depth, nxt, free[(k + 1) % self.depth])
            except StopIteration:
                pending = None
            torch.cuda.current_stream().wait_event(ready)
            yield cur
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream())     # everything the consumer enqueued on this batch
            free[k % self.depth] = done
            k += 1


def get_grad_norm_(parameters, norm_type: float = 2.0) -> torch.Tensor:
    """Global gradient norm (reference misc.py:280-292), one fused multi-tensor reduction instead of a launch per tensor."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad.detach() for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.)
    if norm_type == float("inf"):
        return torch.stack([g.abs().max() for g in grads]).max()
    return torch.linalg.vector_norm(torch.stack(torch._foreach_norm(grads, norm_type)), norm_type)


class NativeScalerWithGradNormCount:
    """Same call contract as the reference's wrapper (misc.py:251-277): scale -> backward -> (unscale, norm | clip,
    step, update) when ``update_grad``; returns the gradient norm or None.

    When ``optimizer`` is a plain ``torch.optim.AdamW`` over the parameters of one of this package's models (the
    optimizer the k-fold scripts build, k_fold_cross_valid_combined_brats.py:168-169) and no clipping is requested,
    unscale + norm + AdamW + scale update run as three kernels over the model's flat buffers
    (engine.FusedAdamW) -- same arithmetic, same optimizer state layout; the loss scale then lives on the device.
    Anything else takes the reference's torch path unchanged."""
    state_dict_key = "amp_scaler"

    def __init__(self):
        self._scaler = torch.amp.GradScaler("cuda")
        self._fused = None
        self.allow_fused = True
        self.allow_sharded = True      # data parallel: dp.ShardedStep instead of all-reduce + replicated AdamW when available
        # True: the fused AdamW pass runs per layer group on its own stream and the next forward waits group by group
        # (engine.FusedAdamW.step).  Only for callers that do not touch parameters / gradients / optimizer state between
        # this call and the next forward.  Off by default: measured on B200 (ViT-B, batch 4) the overlapped pair takes as
        # long as the serial one -- AdamW saturates HBM/L2 and the forward's GEMMs are bound by the same L2 -> SM path.
        self.overlap_optimizer = False

    def _fused_for(self, optimizer, clip_grad, create_graph):
        if not self.allow_fused or clip_grad is not None or create_graph:
            return None
        from ..engine import FusedAdamW
        if not FusedAdamW.supports(optimizer):
            return None
        ref = None
        for group in optimizer.param_groups:       # the first parameter that belongs to one of this package's engines
            for p in group["params"]:
                ref = getattr(p, "_vitae_engine", None)
                if ref is not None:
                    break
            if ref is not None:
                break
        eng = ref() if ref is not None else None
        if eng is None or not eng.flat.still_aliased():
            return None
        fo = eng.fused_optimizer()
        if not fo.bind(optimizer):
            return None
        if self._fused is not fo:      # hand the loss-scale state to the device-side control block
            sd = self._scaler.state_dict() if self._scaler.is_enabled() else {}
            fo.ctl[0] = float(sd.get("scale", 1.0))
            fo.ctl[1] = float(sd.get("_growth_tracker", 0))
            self._fused = fo
        return fo

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True):
        fo = self._fused_for(optimizer, clip_grad, create_graph)
        if fo is not None:
            eng = fo.eng
            # data parallel on one NVLink domain: the step below reduces + updates shard-wise (dp.ShardedStep), so the
            # backward must not all-reduce; it reduces each finished slice into its owner when the step follows
            sharded = fo.sharded() if self.allow_sharded else None
            eng.defer_exchange = None if sharded is None else ("reduce" if update_grad else "skip")
            try:
                (loss * fo.ctl[0]).backward()
            finally:
                eng.defer_exchange = None
            if not update_grad:
                return None
            return fo.step(optimizer, self._scaler, overlap=self.overlap_optimizer)
        if self._fused is not None:    # leaving the fused path: give the scale state back to torch's scaler
            self._scaler.load_state_dict(self.state_dict())
            if getattr(self._fused.eng, "grads_local", False):      # accumulated, not yet exchanged gradients
                self._fused.eng.allreduce_gradients()
            self._fused.sync_state(optimizer)
            self._fused = None
        self._scaler.scale(loss).backward(create_graph=create_graph)
        if not update_grad:
            return None
        self._scaler.unscale_(optimizer)
        if clip_grad is not None:
            assert parameters is not None
            norm = torch.nn.utils.clip_grad_norm_(parameters, clip_grad)
        else:
            norm = get_grad_norm_(parameters)
        self._scaler.step(optimizer)
        self._scaler.update()
        return norm

    def state_dict(self):
        if self._fused is None:
            return self._scaler.state_dict()
        scale, tracker = self._fused.ctl[:2].tolist()
        return {"scale": scale, "growth_factor": self._scaler.get_growth_factor(),
                "backoff_factor": self._scaler.get_backoff_factor(),
                "growth_interval": self._scaler.get_growth_interval(), "_growth_tracker": int(tracker)}

    def load_state_dict(self, state_dict):
        self._scaler.load_state_dict(state_dict)
        if self._fused is not None and state_dict:
            self._fused.ctl[0] = float(state_dict["scale"])
            self._fused.ctl[1] = float(state_dict["_growth_tracker"])


def save_model(args, epoch, model, model_without_ddp, optimizer, loss_scaler):
    """``<args.output_dir>/checkpoint-<epoch>.pth`` = {model, optimizer, epoch, scaler, args}, written by rank 0
    (reference misc.py:295-310).  ``optimizer.state_dict()`` is complete at any time: the fused AdamW keeps its moments
    in views the optimizer's state already points at and refreshes the per-parameter ``step`` through a state-dict
    pre-hook (engine.FusedAdamW.bind).  ``loss_scaler=None`` is the reference's DeepSpeed branch."""
    if loss_scaler is None:
        model.save_checkpoint(save_dir=args.output_dir, tag="checkpoint-%s" % str(epoch), client_state={"epoch": epoch})
        return
    path = Path(args.output_dir) / ("checkpoint-%s.pth" % str(epoch))
    save_on_master({"model": model_without_ddp.state_dict(), "optimizer": optimizer.state_dict(), "epoch": epoch,
                    "scaler": loss_scaler.state_dict(), "args": args}, path)


def load_model(args, model_without_ddp, optimizer, loss_scaler):
    """Resumes model (+ optimizer and loss scale unless ``args.eval``) from ``args.resume`` (path or https URL); a falsy
    ``args.resume`` is a no-op (reference misc.py:313-329).  Checkpoints hold an ``argparse.Namespace`` (``args``), hence
    ``weights_only=False`` -- the reference's plain ``torch.load`` predates that default."""
    if not getattr(args, "resume", None):
        return
    if args.resume.startswith("https"):
        checkpoint = torch.hub.load_state_dict_from_url(args.resume, map_location="cpu", check_hash=True)
    else:
        checkpoint = torch.load(args.resume, map_location="cpu", weights_only=False)
    model_without_ddp.load_state_dict(checkpoint["model"])
    print("Resume checkpoint %s" % args.resume)
    if "optimizer" in checkpoint and "epoch" in checkpoint and not getattr(args, "eval", False):
        optimizer.load_state_dict(checkpoint["optimizer"])
        if "scaler" in checkpoint:
            loss_scaler.load_state_dict(checkpoint["scaler"])
        print("With optim & sched!")


def all_reduce_mean(x):
    """Mean over ranks of a python scalar or 0-dim tensor (reference misc.py:332-340)."""
    world = get_world_size()
    if world == 1:
        return x
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.as_tensor(x, dtype=torch.float32).detach().to(dev).clone()
    dist.all_reduce(t)
    return (t / world).item()


def add_weight_decay(model, weight_decay=1e-5, skip_list=()):
    """timm 0.5.4 ``optim_factory.add_weight_decay`` semantics, as called at k_fold_cross_valid_combined_brats.py:168:
    frozen parameters are skipped; 1-D tensors, ``.bias`` and names in ``skip_list`` get no decay; everything else
    (including the 3-D cls / mask tokens) is decayed."""
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if (p.ndim == 1 or name.endswith(".bias") or name in skip_list) else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.}, {"params": decay, "weight_decay": weight_decay}]
