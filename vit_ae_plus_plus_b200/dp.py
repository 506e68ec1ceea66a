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


# ---------------------------------------------------------------------------------------------------------------------
# Sharded step over NVLink peer memory
# ---------------------------------------------------------------------------------------------------------------------
GRANULE_SHIFT = 16          # ownership granule: 2^16 elements (256 KB of fp32); granule q belongs to rank q % world
# Symmetric allocations (and the rendezvous handles that map them into the peers) must not be destroyed at an arbitrary
# moment: unmapping while some stream is being captured into a CUDA graph aborts the process.  Everything allocated here is
# therefore kept alive in this registry together with a weak reference to its owner (the FlatParams object), and entries of
# dead owners are dropped at the next allocation -- a point where this package is certainly not capturing.
_symmetric_ptrs = set()
_registry = []              # [weakref to the owner, [tensors / handles]]


def _purge_registry() -> None:
    dead = [e for e in _registry if e[0]() is None]
    if not dead:
        return
    torch.cuda.synchronize()
    for e in dead:
        for obj in e[1]:
            if isinstance(obj, torch.Tensor):
                _symmetric_ptrs.discard(obj.data_ptr())
        _registry.remove(e)
    del dead


def keep_alive(owner, objects) -> None:
    """Ties symmetric tensors / rendezvous handles to ``owner``: they live at least as long as it does and are released
    at a safe point afterwards."""
    import weakref
    for e in _registry:
        if e[0]() is owner:
            e[1].extend(objects)
            return
    _registry.append([weakref.ref(owner), list(objects)])


def sharded_enabled() -> bool:
    """The sharded step needs every rank of the job on one NVLink domain: an NCCL process group of 2..8 ranks on one node
    (torchrun's LOCAL_WORLD_SIZE == WORLD_SIZE).  VITAE_DP_SHARDED=0 keeps the all-reduce + replicated AdamW path."""
    if os.environ.get("VITAE_DP_SHARDED", "1") == "0":
        return False
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend() != "nccl":
        return False
    w = dist.get_world_size()
    if w < 2 or w > 8:
        return False
    return os.environ.get("LOCAL_WORLD_SIZE", str(w)) == str(w)


def alloc_flat(n: int, dtype: torch.dtype, device, owner=None) -> torch.Tensor:
    """Zero-filled flat buffer.  With the sharded step available it comes from torch's symmetric-memory allocator
    (cuMemCreate + a mapping into every peer after ShardedStep's rendezvous: plumbing, like torch.empty) and is tied to
    ``owner`` (keep_alive); else torch.zeros."""
    device = torch.device(device)
    if device.type == "cuda" and sharded_enabled() and owner is not None:
        try:
            import torch.distributed._symmetric_memory as symm_mem
            _purge_registry()
            t = symm_mem.empty(n, dtype=dtype, device=device)
            t.zero_()
            _symmetric_ptrs.add(t.data_ptr())
            keep_alive(owner, [t])
            return t
        except Exception as e:      # noqa: BLE001 -- allocator not usable on this system: the NCCL path still works
            import warnings
            warnings.warn(f"symmetric allocation failed ({type(e).__name__}: {e}); using the all-reduce path")
    return torch.zeros(n, dtype=dtype, device=device)


def is_symmetric(t: torch.Tensor) -> bool:
    return t.data_ptr() in _symmetric_ptrs


class ShardedStep:
    """Data-parallel optimizer step without an all-reduce and without replicated optimizer work.

    all-reduce path:  backward || all-reduce(grads)  ->  every rank: AdamW over ALL parameters (0.69 ms for ViT-B)
    sharded path:     backward || reduce of the OWNED part of each finished slice (vitae_dp_reduce_shard pulls the peers'
                      copies over NVLink)  ->  barrier  ->  AdamW over the owned 1/world of the parameters with the
                      all-gather fused in (vitae_adamw_shard stores the new bf16 shadow into every rank's buffer, fp32 only
                      for the small tensors the kernels read in fp32)  ->  barrier.
    NVLink bytes per rank and step: (world-1)/world * (4 + 2) bytes per parameter, against 2 * (world-1)/world * 4 for an
    fp32 all-reduce; optimizer HBM traffic drops by the factor world.

    What differs from a replicated step, by design: between steps the fp32 master of the large matrices and the Adam
    moments are current only on their owner.  ``sync_master()`` / ``sync_moments()`` pull the missing parts from the owners
    (peer reads, no collective; the model's / optimizer's ``state_dict()`` and ``FlatParams.refresh_shadow`` call them), and
    ``.grad`` holds the mean gradient only in the owned part after ``backward``.

    Cross-rank ordering uses the signal-pad barrier of the symmetric allocation (torch plumbing, one small kernel)."""

    PART_CAP = 148 * 4            # partial sums per slice (vitae_dp_reduce_shard_blocks' cap)
    MAX_SLICES = 24

    def __init__(self, flat, m: torch.Tensor, v: torch.Tensor, group_map: torch.Tensor):
        import torch.distributed._symmetric_memory as symm_mem
        from . import ops
        self.ops = ops
        self.flat, self.m, self.v, self.group_map = flat, m, v, group_map
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        dev = flat.g32.device
        # per-slice partial sums of squares, then (last 64 floats) their total: the one value the peers read
        self.part = symm_mem.empty(self.MAX_SLICES * self.PART_CAP + 64, dtype=torch.float32, device=dev)
        self.part.zero_()
        self.part_total = self.part[self.MAX_SLICES * self.PART_CAP:]
        group = dist.group.WORLD
        self.h = {name: symm_mem.rendezvous(t, group) for name, t in
                  (("g32", flat.g32), ("p32", flat.p32), ("p16", flat.p16), ("m", m), ("v", v), ("part", self.part))}
        keep_alive(flat, [self.part, *self.h.values()])
        self.peers = {name: ops.peer_table(h.buffer_ptrs) for name, h in self.h.items()}
        self.peers["part_total"] = ops.peer_table([p + 4 * self.MAX_SLICES * self.PART_CAP for p in self.h["part"].buffer_ptrs])
        # chunks whose fp32 master every rank needs: everything but the large matrices (only their bf16 shadow is read)
        wide = torch.ones(flat.total // 64, dtype=torch.uint8)
        for n in flat.order:
            o, k, shp = flat.offsets[n]
            if len(shp) >= 2 and k >= (1 << 16):
                wide[o // 64:(o + k + 63) // 64] = 0
        self.f32_chunk = wide.to(dev)
        self.comm = torch.cuda.Stream(device=dev, priority=-1)
        self.overlap_blocks = int(os.environ.get("VITAE_DP_REDUCE_BLOCKS", "96"))   # grid of a reduce that runs beside the backward
        self.gather_blocks = int(os.environ.get("VITAE_DP_GATHER_BLOCKS", "296"))   # grid of an update that runs beside the forward
        # opt-in: on 2 GPUs the resident step is unchanged (4.085 vs 4.098 ms) and the end-to-end step gets slower and noisy
        # (27 more enqueues per step on a host that also feeds the copy stream; profiles/r02zd_*); not measured at 8
        self.overlap_gather = os.environ.get("VITAE_DP_OVERLAP_GATHER", "0") == "1"
        self.slices: List[Tuple[int, int]] = []      # slices reduced since the last step, in order
        self.cursor = 0                              # partial sums written so far (PART_CAP per slice)
        self.master_stale = False
        self.moments_stale = False
        self.steps = 0

    # ---- reduce ----------------------------------------------------------------------------------------------------
    def begin(self) -> None:
        """Start of a backward whose gradients this object will reduce."""
        self.slices, self.cursor = [], 0

    def launch(self, t: torch.Tensor, overlap: bool = True, final: bool = False) -> None:
        """Gradient slice ``t`` (a view of the flat gradient buffer) is final on the current stream: reduce the owned part.
        overlap=True: on the communication stream with a small grid (the backward keeps running on the current stream);
        final=True: nothing runs beside it any more, full grid."""
        a = (t.data_ptr() - self.flat.g32.data_ptr()) // 4
        b = a + t.numel()
        k = len(self.slices)
        if k >= self.MAX_SLICES:
            raise RuntimeError("ShardedStep: too many gradient slices")
        self.slices.append((a, b))
        part = self.part[k * self.PART_CAP:(k + 1) * self.PART_CAP]
        main = torch.cuda.current_stream()
        if overlap:
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ev)
                self._reduce(a, b, part, k, 0 if final else self.overlap_blocks)
        else:
            self._reduce(a, b, part, k, 0)

    def _reduce(self, a: int, b: int, part: torch.Tensor, k: int, max_blocks: int) -> None:
        part.zero_()                                  # ranks that own nothing of a slice contribute zeros
        self.h["g32"].barrier(channel=k)              # every rank's copy of this slice is final
        self.ops.dp_reduce_shard(self.peers["g32"], self.world, self.rank, a, b, GRANULE_SHIFT, 1.0 / self.world, part,
                                 max_blocks)

    def wait(self) -> None:
        torch.cuda.current_stream().wait_stream(self.comm)

    def reduce_all(self) -> None:
        """Not overlapped: the whole gradient buffer at once, on the current stream."""
        self.begin()
        self.launch(self.flat.g32, overlap=False)

    # ---- step ------------------------------------------------------------------------------------------------------
    def step(self, ctl: torch.Tensor, rows, growth_factor: float, backoff_factor: float, growth_interval: int,
             use_scaler: bool, overlap=None) -> bool:
        """GradScaler.unscale_/update + AdamW on the gradients reduced since begin(); on return (stream order) every rank's
        bf16 shadow and small fp32 tensors are current.

        ``overlap`` = (stream, [(start, end) per parameter group in FORWARD order], [event per group]): the update + all-gather
        is issued group by group on ``stream``, each followed by its own cross-rank barrier and event; the next forward waits
        for a group right before its first use (engine._need), so the all-gather of the later layers -- NVLink-bound, not
        HBM-bound -- hides behind the forward of the earlier ones.  Returns True when issued that way (the caller's stream is
        then NOT ordered after the update until engine.wait_params())."""
        if not self.slices:
            self.reduce_all()
        covered = sum(b - a for a, b in self.slices)
        if covered != self.flat.total:
            raise RuntimeError(f"ShardedStep: reduced slices cover {covered} of {self.flat.total} gradient elements")
        ops, W, r = self.ops, self.world, self.rank
        ops.sum_partials(self.part, len(self.slices) * self.PART_CAP, self.part_total)
        self.h["part"].barrier(channel=self.MAX_SLICES)          # every rank's total is written
        ops.optim_finalize_peers(self.peers["part_total"], W, 1, ctl, growth_factor, backoff_factor, growth_interval,
                                 use_scaler)
        if overlap is None:
            # one launch: ownership does not depend on how the backward sliced the buffer
            ops.adamw_shard(self.peers["p32"], self.peers["p16"], W, r, 0, self.flat.total, GRANULE_SHIFT, self.flat.g32,
                            self.m, self.v, self.group_map, self.f32_chunk, rows, ctl)
            self.h["p16"].barrier(channel=self.MAX_SLICES + 1)       # every owner's stores have landed everywhere
        else:
            stream, ranges, events = overlap
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(stream):
                stream.wait_event(ev)
                for g, (a, b) in enumerate(ranges):
                    ops.adamw_shard(self.peers["p32"], self.peers["p16"], W, r, a, b, GRANULE_SHIFT, self.flat.g32, self.m,
                                    self.v, self.group_map, self.f32_chunk, rows, ctl, max_blocks=self.gather_blocks)
                    self.h["p16"].barrier(channel=self.MAX_SLICES + 2 + g)
                    events[g].record(stream)
        self.slices, self.cursor = [], 0
        self.master_stale = self.moments_stale = True
        self.steps += 1
        return overlap is not None

    # ---- on-demand completion of the replicas ----------------------------------------------------------------------------
    def _pull(self, name: str, local: torch.Tensor) -> None:
        """local[granules owned by peers] := the owner's copy (peer reads over NVLink; the owners are not involved -- their
        copies cannot change before this rank enters the next step's barriers)."""
        h, W, G = self.h[name], self.world, 1 << GRANULE_SHIFT
        total = local.numel()
        nfull = total // G
        for src in range(W):
            if src == self.rank:
                continue
            remote = h.get_buffer(src, (total,), local.dtype, 0)
            if nfull > src:
                local[:nfull * G].view(nfull, G)[src::W].copy_(remote[:nfull * G].view(nfull, G)[src::W])
            if total > nfull * G and nfull % W == src:      # the ragged last granule
                local[nfull * G:].copy_(remote[nfull * G:])

    def sync_master(self) -> None:
        """fp32 master of the large matrices: fetch the parts other ranks own (they were updated there, only their bf16
        shadow came here)."""
        if self.master_stale:
            self._pull("p32", self.flat.p32)
            self.master_stale = False

    def sync_moments(self) -> None:
        if self.moments_stale:
            self._pull("m", self.m)
            self._pull("v", self.v)
            self.moments_stale = False


def replica_check(model, volumes, noises, opt_steps: int = 4) -> dict:
    """Self-check of the data-parallel path on the live process group (tests/dp_check.py under torchrun; bench.py runs it
    before the timed region of every N > 1 run and prints the result in its JSON line):
      1. the flat parameter buffer is identical on every rank after construction (ranks seed differently; rank 0 wins);
      2. ``backward`` with the staged, overlapped gradient exchange leaves in every rank's ``.grad`` the mean over ranks
         of the local gradients (compared with a ``no_sync`` backward + one explicit all-reduce), three times: eager
         launches, graph capture, graph replay;
      3. after ``opt_steps`` fused GradScaler + AdamW steps on rank-local data the replicas are still bit-identical (bf16
         shadow and, after ``sync_master``, fp32 master); when those steps ran sharded (ShardedStep) the same steps are
         repeated through all-reduce + replicated AdamW and the two results compared.
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
    p0 = eng.flat.p32.clone()

    def run_steps(allow_sharded: bool):
        opt = torch.optim.AdamW(misc.add_weight_decay(model, 0.05), lr=1e-3, betas=(0.9, 0.95))
        scaler = misc.NativeScalerWithGradNormCount()
        scaler.allow_sharded = allow_sharded
        for i in range(opt_steps):
            losses = model(volumes[i % len(volumes)], noise=noises[i % len(noises)])[0]
            scaler(losses[0], opt, parameters=model.parameters(), update_grad=True)
            opt.zero_grad()
        return scaler, losses[0].item(), opt

    scaler, final_loss, opt = run_steps(True)
    sh = eng.flat.sharded
    out["sharded"] = bool(sh is not None and sh.steps > 0)
    eng.wait_params()                                                    # (an overlapped update may still be in flight)
    out["shadow_spread_after_steps"] = spread(eng.flat.p16.float())      # what the kernels compute with
    eng.sync_master()                                                    # sharded: complete the fp32 master first
    out["param_spread_after_steps"] = spread(eng.flat.p32)
    out["fused_optimizer"] = scaler._fused is not None
    out["moment_spread_after_state_dict"] = 0.0
    if scaler._fused is not None:
        opt.state_dict()                   # its pre-hook completes the moments (sharded: peer reads from the owners)
        out["moment_spread_after_state_dict"] = max(spread(scaler._fused.m), spread(scaler._fused.v))
    out["final_loss"] = final_loss
    out["exchange_dtype"] = exchange_dtype()
    out["sharded_vs_allreduce"] = None
    if out["sharded"]:
        # the same steps from the same start through all-reduce + replicated AdamW: mean |difference| of the parameters
        # relative to the mean |update| (the reductions add in a different order, Adam normalises: not bit-identical)
        p_sh = eng.flat.p32.clone()
        with torch.no_grad():
            eng.flat.p32.copy_(p0)
        run_steps(False)
        eng.wait_params()
        upd = (eng.flat.p32 - p0).abs().mean().item()
        out["sharded_vs_allreduce"] = (p_sh - eng.flat.p32).abs().mean().item() / (upd + 1e-30)
        out["param_spread_after_steps"] = max(out["param_spread_after_steps"], spread(eng.flat.p32))
    # fp32 exchange reproduces the explicit fp32 mean; bf16 exchange rounds every rank's slice to bf16 first (2^-9 relative)
    tol = 1e-6 if out["exchange_dtype"] == "fp32" else 1e-2
    out["ok"] = (out["param_spread_after_broadcast"] == 0.0 and out["grad_spread_after_exchange"] == 0.0
                 and out["param_spread_after_steps"] == 0.0 and out["shadow_spread_after_steps"] == 0.0
                 and out["moment_spread_after_state_dict"] == 0.0
                 and out["exchange_rel_err"] <= tol
                 and (out["sharded_vs_allreduce"] is None or out["sharded_vs_allreduce"] <= 1e-2))
    return out
