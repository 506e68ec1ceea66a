"""Data-parallel plumbing of the training step: volumes are sharded over ranks, every rank
holds a full replica, and the only exchange is one mean all-reduce of the flat gradient buffer per optimizer step
(plus one broadcast of the flat parameter buffer at construction: the k-fold scripts seed ranks differently,
k_fold_cross_valid_combined_brats.py:87, and never wrap the model in DDP, :154).

The flat buffers are laid out in the order gradients complete during backward (engine.backward_param_order), so a
bucket is a contiguous slice and buckets become ready front to back.  ``torch.distributed`` is the transport (NCCL over
NVLink on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def bucket_slices(tensor_offsets: Sequence[Tuple[int, int]], total: int, bucket_elems: int) -> List[Tuple[int, int]]:
    """Splits [0, total) into contiguous slices of about ``bucket_elems`` elements that end on tensor boundaries
    (``tensor_offsets`` = (offset, numel) per tensor in layout order; a tensor larger than a bucket gets its own)."""
    if total <= 0:
        return []
    cuts, start = [], 0
    ends = [o + k for o, k in tensor_offsets]
    for i, end in enumerate(ends):
        last = i == len(ends) - 1
        if last:
            end = total
        if end - start >= bucket_elems or last:
            cuts.append((start, end))
            start = end
    return cuts


def broadcast_flat(t: torch.Tensor, src: int = 0) -> None:
    if world_size() > 1:
        dist.broadcast(t, src=src)


def allreduce_mean_(t: torch.Tensor, async_op: bool = False):
    """In-place mean over ranks.  NCCL averages in the collective; gloo has no AVG, so sum then scale."""
    w = world_size()
    if w == 1:
        return None
    if dist.get_backend() == "nccl":
        return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
    work = dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=False)
    t.div_(w)
    return work


def allreduce_mean_bucketed_(flat: torch.Tensor, slices: Sequence[Tuple[int, int]]) -> None:
    for a, b in slices:
        allreduce_mean_(flat[a:b])


def exchange_dtype() -> str:
    """'fp32' (default) or 'bf16' (VITAE_GRAD_EXCHANGE=bf16, NCCL only): the element type in which gradient slices cross
    NVLink.  bf16 halves the bytes of the one exchange the step has (527 -> 263 MB per step for ViT-B): every rank's slice
    is rounded to bf16 before the reduction and the mean is widened back into the fp32 gradient buffer; optimizer and
    loss-scale arithmetic stay fp32.  Measured on 2 x B200 (profiles/r02i_bench2_*.json) the two staging copies cost more
    than the halved transfer saves (5.06 vs 4.74 ms per step), so it is opt-in for larger rank counts."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend() != "nccl":
        return "fp32"
    return "bf16" if os.environ.get("VITAE_GRAD_EXCHANGE", "fp32").lower() in ("bf16", "bfloat16") else "fp32"


class GradReducer:
    """Mean all-reduce of gradient slices, issued asynchronously while the caller keeps enqueueing the next backward
    stage; ``wait()`` orders the current stream after all of them (no host block with NCCL).

    fp32 exchange: the slice's collective runs on the process group's own stream, ordered after everything enqueued on
    the current stream so far.  bf16 exchange (``staging`` = a bf16 buffer shaped like the flat gradient buffer, ``base`` =
    the flat fp32 buffer the slices are views of): narrow -> all-reduce -> widen run on a communication stream of their
    own (vitae_cast_f32_to_bf16 / vitae_cast_bf16_to_f32 with a capped grid: they share the SMs with the backward)."""

    def __init__(self, base: torch.Tensor = None, staging: torch.Tensor = None, stream=None):
        self.pending = []
        self.base, self.staging, self.stream = base, staging, stream

    def launch(self, t: torch.Tensor) -> None:
        w = world_size()
        if w == 1 or t.numel() == 0:
            return
        if dist.get_backend() != "nccl":      # gloo: no AVG
            self.pending.append((dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True), t))
            return
        if self.staging is None:
            self.pending.append((dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=True), None))
            return
        from . import ops
        off = (t.data_ptr() - self.base.data_ptr()) // 4
        h = self.staging[off:off + t.numel()]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            ops.cast_f32_to_bf16(t, h)
            work = dist.all_reduce(h, op=dist.ReduceOp.AVG, async_op=True)
            work.wait()                       # the communication stream waits for NCCL's
            ops.cast_bf16_to_f32(h, t)
        self.pending.append((None, None))

    def wait(self) -> None:
        w = world_size()
        for work, t in self.pending:
            if work is not None:
                work.wait()
                if t is not None:
                    t.div_(w)
        if self.staging is not None and self.pending:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.pending = []


def replica_check(model, volumes, noises, opt_steps: int = 4) -> dict:
    """Self-check of the data-parallel path on the live process group (tests/dp_check.py under torchrun; bench.py runs it
    before the timed region of every N > 1 run and prints the result in its JSON line):
      1. the flat parameter buffer is identical on every rank after construction (ranks seed differently; rank 0 wins);
      2. ``backward`` with the staged, overlapped gradient exchange leaves in every rank's ``.grad`` the mean over ranks
         of the local gradients (compared with a ``no_sync`` backward + one explicit all-reduce), three times: eager
         launches, graph capture, graph replay;
      3. after ``opt_steps`` fused GradScaler + AdamW steps on rank-local data the replicas are still bit-identical.
    ``volumes``: rank-local device batches; ``noises``: rank-local mask noise, one per step.  Restores nothing: call it
    on a throw-away model.  Returns the measured spreads / errors (all must be 0 except ``exchange_rel_err`` <= 1e-6)."""
    from .utils import misc
    eng = model.engine()

    def spread(t):
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return (hi - lo).abs().max().item()

    out = {"world": world_size(), "param_spread_after_broadcast": spread(eng.flat.p32), "exchange_rel_err": 0.0,
           "grad_spread_after_exchange": 0.0}
    for rep in range(3):
        with model.no_sync():
            losses = model(volumes[0], noise=noises[0])[0]
            losses[0].backward()
        ref = eng.flat.g32.clone()
        allreduce_mean_(ref)
        for p in model.parameters():
            p.grad = None
        losses = model(volumes[0], noise=noises[0])[0]
        losses[0].backward()
        got = eng.flat.g32.clone()
        for p in model.parameters():
            p.grad = None
        out["exchange_rel_err"] = max(out["exchange_rel_err"],
                                      (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-30))
        out["grad_spread_after_exchange"] = max(out["grad_spread_after_exchange"], spread(got))
    opt = torch.optim.AdamW(misc.add_weight_decay(model, 0.05), lr=1e-3, betas=(0.9, 0.95))
    scaler = misc.NativeScalerWithGradNormCount()
    for i in range(opt_steps):
        losses = model(volumes[i % len(volumes)], noise=noises[i % len(noises)])[0]
        scaler(losses[0], opt, parameters=model.parameters(), update_grad=True)
        opt.zero_grad()
    out["param_spread_after_steps"] = spread(eng.flat.p32)
    out["fused_optimizer"] = scaler._fused is not None
    out["final_loss"] = losses[0].item()
    out["exchange_dtype"] = exchange_dtype()
    # fp32 exchange reproduces the explicit fp32 mean; bf16 exchange rounds every rank's slice to bf16 first (2^-9 relative)
    tol = 1e-6 if out["exchange_dtype"] == "fp32" else 1e-2
    out["ok"] = (out["param_spread_after_broadcast"] == 0.0 and out["grad_spread_after_exchange"] == 0.0
                 and out["param_spread_after_steps"] == 0.0 and out["exchange_rel_err"] <= tol)
    return out
