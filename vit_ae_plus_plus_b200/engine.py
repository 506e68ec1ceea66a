"""Kernel orchestration of the 3D ViT masked-autoencoder training step (forward + hand-derived backward).

This is the host side of the hot path: it owns the flat parameter / gradient buffers and a per-shape activation arena,
and enqueues the sm_100a kernels of libvitae_b200.so (through ops.py -> C ABI) in the order the reference's
``MaskedAutoencoderViT.forward`` (model/vit_autoenc.py:234-238) and ``loss.backward()`` (utils/misc.py:258) imply.
Nothing here computes with torch: torch provides device memory, streams and (for N>1) torch.distributed.

Data layout in HBM
  parameters   one flat fp32 buffer (master copy; the nn.Parameters are views into it) + one flat bf16 shadow with the
               same offsets (GEMM operands) + one flat fp32 gradient buffer (wgrad epilogues write into it directly).
               Tensors are ordered by the time their gradient completes in backward (decoder_pred first, patch_embed
               last) so that data-parallel buckets are contiguous slices that become ready front to back.
  tokens       residual stream fp32 [B*N, D] row-major, sample-major (row = b*N + token, cls token = row b*N);
               GEMM operands (LayerNorm outputs, qkv, attention output, MLP hidden) bf16 with the same row order.
  volume       the caller's fp32 NCDHW tensor is read in place by the patch gather and by the loss kernels
               (no patchify copy, model/vit_autoenc.py:100-113).

Execution
  one CUDA graph per (shape, input address) for the forward kernels and one for the backward kernels: ~500 kernel
  nodes are replayed with a single launch each (a ViT-B step is ~2 ms of device time; launching its kernels one by one
  from Python costs ~12 ms).  Inside the backward graph the weight-gradient GEMMs run on a side stream and the per-block
  column reductions (bias sums, LayerNorm affine gradients) on a third (they are not on the dgrad critical path, which
  is captured at a higher stream priority), with per-buffer events guarding every read-after-write / write-after-read
  pair between the lanes and gradient rings deep enough that the chain never waits for a lane (_Lanes, ring_depths).
  The edge-map target branch (input only) runs on a background stream underneath the forward.

Mixed precision: bf16 tensor-core operands, fp32 accumulation, fp32 residual stream / LayerNorm statistics /
loss / gradients of parameters (BASELINE.json north_star tolerance for this mode: 1e-2 relative).
"""
from __future__ import annotations

import gc
import math
import os
import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import dp, ops

_F32, _BF16, _I32 = torch.float32, torch.bfloat16, torch.int32
_ALIGN = 64  # elements; keeps every tensor 128-byte aligned in the bf16 shadow (TMA needs 16)


@dataclass
class StackSpec:
    prefix: str      # 'blocks' or 'decoder_blocks'
    dim: int
    heads: int
    hidden: int
    depth: int

    @property
    def head_dim(self) -> int:
        return self.dim // self.heads


def block_param_names(prefix: str, i: int) -> List[str]:
    """Parameter names of one transformer block in the order their gradients complete in backward."""
    b = f"{prefix}.{i}"
    return [f"{b}.mlp.fc2.weight", f"{b}.mlp.fc2.bias", f"{b}.mlp.fc1.weight", f"{b}.mlp.fc1.bias",
            f"{b}.norm2.weight", f"{b}.norm2.bias", f"{b}.attn.proj.weight", f"{b}.attn.proj.bias",
            f"{b}.attn.qkv.weight", f"{b}.attn.qkv.bias", f"{b}.norm1.weight", f"{b}.norm1.bias"]


PREDICTOR_PARAMS = ["predictor.3.weight", "predictor.3.bias", "predictor.1.weight", "predictor.1.bias", "predictor.0.weight"]


def backward_param_order(depth: int, decoder_depth: int, predictor: bool = False) -> List[str]:
    # the contrastive predictor (model/vit_autoenc.py:263-268) sits downstream of everything else: its gradients come first
    names = list(PREDICTOR_PARAMS) if predictor else []
    names += ["decoder_pred.weight", "decoder_pred.bias", "decoder_norm.weight", "decoder_norm.bias"]
    for i in reversed(range(decoder_depth)):
        names += block_param_names("decoder_blocks", i)
    names += ["mask_token", "decoder_embed.weight", "decoder_embed.bias", "norm.weight", "norm.bias"]
    for i in reversed(range(depth)):
        names += block_param_names("blocks", i)
    names += ["cls_token", "patch_embed.proj.weight", "patch_embed.proj.bias"]
    return names


class FlatParams:
    """Flat fp32 master / bf16 shadow / fp32 gradient buffers; the module's nn.Parameters become views of ``p32``."""

    def __init__(self, named_params: Dict[str, torch.nn.Parameter], order: List[str], device: torch.device):
        assert sorted(order) == sorted(named_params.keys()), "parameter set does not match the engine's layout"
        self.order = order
        self.offsets: Dict[str, Tuple[int, int, torch.Size]] = {}
        off = 0
        for n in order:
            p = named_params[n]
            self.offsets[n] = (off, p.numel(), p.shape)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = off
        # (data parallel on one NVLink domain: symmetric allocations, see dp.ShardedStep; plain torch.zeros otherwise)
        self.p32 = dp.alloc_flat(off, _F32, device, owner=self)
        self.p16 = dp.alloc_flat(off, _BF16, device, owner=self)
        self.g32 = dp.alloc_flat(off, _F32, device, owner=self)
        self.sharded = None           # dp.ShardedStep once the fused optimizer has set it up
        self.params = named_params
        self.v32: Dict[str, torch.Tensor] = {}
        self.v16: Dict[str, torch.Tensor] = {}
        self.vg: Dict[str, torch.Tensor] = {}
        with torch.no_grad():
            for n in order:
                o, k, shp = self.offsets[n]
                self.v32[n] = self.p32[o:o + k].view(shp)
                self.v16[n] = self.p16[o:o + k].view(shp)
                self.vg[n] = self.g32[o:o + k].view(shp)
                self.v32[n].copy_(named_params[n].data)
                named_params[n].data = self.v32[n]
        # one-record table for the cast kernel (the flat buffer is a single contiguous tensor)
        self.cast_table = torch.tensor([[self.p32.data_ptr(), 0, self.total]], dtype=torch.int64, device=device)
        self.shadow_version = -1      # parameter version stamp the bf16 shadow was produced from (-1: never)
        self._plist = [named_params[n] for n in order]
        self.overwrite_grads = False  # set by mark_grads_consumed(): next backward overwrites instead of accumulating

    def still_aliased(self) -> bool:
        n = self.order[0]
        return self.params[n].data_ptr() == self.v32[n].data_ptr() and self.params[n].device == self.p32.device

    def _version(self) -> int:
        # nn.Parameter.data = view gives every parameter its own version counter: in-place writes through torch
        # (optimizer.step, load_state_dict, init) bump the parameter's, not the flat buffer's
        return sum(p._version for p in self._plist) + self.p32._version

    def refresh_shadow(self) -> None:
        """fp32 master -> bf16 shadow (one launch).  Skipped while nothing has written the master through torch since
        the shadow was produced (the fused optimizer writes both copies itself)."""
        v = self._version()
        if self.shadow_version != v:
            if self.sharded is not None:      # something wrote the master through torch: complete it first (the parts other
                self.sharded.sync_master()    # ranks own are stale here), then every rank recasts its full copy
            ops.cast_params_bf16(self.cast_table, 1, self.p16, self.total)
            self.shadow_version = v

    def stamp_shadow(self) -> None:
        self.shadow_version = self._version()

    def invalidate_shadow(self) -> None:
        """Forces the next forward to re-cast the bf16 shadow.  Needed after writes that bump no version counter:
        ``param.data.copy_()/mul_()`` (EMA, manual re-initialisation) go around autograd's bookkeeping."""
        self.shadow_version = -1

    def grads_alias(self, thorough: bool = False) -> Optional[bool]:
        """True: every .grad is our view (accumulate in place); False: every .grad is None; None: mixed / foreign.
        The per-step check looks at three sentinel tensors; ``thorough`` walks all of them."""
        state = None
        names = self.order if thorough or len(self.order) < 4 else (self.order[0], self.order[len(self.order) // 2],
                                                                    self.order[-1])
        for n in names:
            g = self.params[n].grad
            s = False if g is None else (True if g.data_ptr() == self.vg[n].data_ptr() else None)
            if s is None:
                return None
            if state is None:
                state = s
            elif state != s:
                return None
        return state


class _Lanes:
    """Main lane = torch's current stream; side lanes = work off the critical path: lane 0 the weight-gradient GEMMs, lane 1
    the per-block column reductions (bias sums, LayerNorm affine gradients: latency-bound two-phase kernels that would
    otherwise sit between the weight-gradient GEMMs).  Hazards between the lanes are ordered with events: ``side()``
    starts after everything enqueued on main so far and remembers which buffers it reads; ``before_write()`` makes main
    wait for the last side readers of a buffer."""

    def __init__(self, device):
        # BELOW the main lane (graphs are captured at priority -1, MAEEngine._capture_stream): the dependency chain gets the
        # SMs it asks for and the side lanes fill what it leaves idle.  (While the main-lane kernels released their
        # dependents at their start, the waiting dependents held the SMs and the side lanes needed the higher priority;
        # with the late release, ptx.cuh, the order flipped: 4.35 -> 4.17 ms per step.)
        self.streams = [torch.cuda.Stream(device=device, priority=int(os.environ.get("VITAE_SIDE_PRIORITY", "0"))),
                        torch.cuda.Stream(device=device, priority=int(os.environ.get("VITAE_REDUCE_PRIORITY", "0")))]
        self.stream = self.streams[0]
        self.readers: Dict[object, Dict[int, torch.cuda.Event]] = {}
        self.dirty = [False, False]
        self.extra_dirty: Optional[torch.cuda.Stream] = None   # another forked stream to join (L2 prefetches)
        self.forked: List[torch.cuda.Stream] = []             # further streams to join (gradient-norm partials)

    def side(self, fn, reads=(), lane: int = 0) -> None:
        main = torch.cuda.current_stream()
        st = self.streams[lane]
        ev = torch.cuda.Event()
        ev.record(main)
        st.wait_event(ev)
        with torch.cuda.stream(st):
            fn()
        done = torch.cuda.Event()
        done.record(st)
        for k in reads:
            self.readers.setdefault(k, {})[lane] = done
        self.dirty[lane] = True

    def before_write(self, *keys) -> None:
        main = torch.cuda.current_stream()
        for k in keys:
            for ev in self.readers.pop(k, {}).values():
                main.wait_event(ev)

    def join(self) -> None:
        if self.extra_dirty is not None:
            ev = torch.cuda.Event()
            ev.record(self.extra_dirty)
            torch.cuda.current_stream().wait_event(ev)
            self.extra_dirty = None
        for lane, st in enumerate(self.streams):
            if self.dirty[lane]:
                ev = torch.cuda.Event()
                ev.record(st)
                torch.cuda.current_stream().wait_event(ev)
                self.dirty[lane] = False
        for st in self.forked:
            ev = torch.cuda.Event()
            ev.record(st)
            torch.cuda.current_stream().wait_event(ev)
        self.forked = []
        self.readers.clear()

    def after_all(self, st: torch.cuda.Stream, fn) -> None:
        """Runs ``fn`` on stream ``st`` after everything enqueued so far on the main lane AND the side lanes; ``join``
        orders the main lane after it."""
        for src in [torch.cuda.current_stream()] + [s for lane, s in enumerate(self.streams) if self.dirty[lane]]:
            ev = torch.cuda.Event()
            ev.record(src)
            st.wait_event(ev)
        with torch.cuda.stream(st):
            fn()
        if st not in self.forked:
            self.forked.append(st)


class _GraphSlot:
    """eager warm-up on first use, capture on the second, replay afterwards"""
    __slots__ = ("graph", "launches", "calls")

    def __init__(self):
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches = 0
        self.calls = 0


# depth of the buffer rings shared between the backward lanes (MAEPlan): residual-stream gradients and LayerNorm inputs
# (two per transformer block) / hidden and qkv gradients (one per block).  The main lane waits before it overwrites a
# slot a side lane still reads, so the depth is how many blocks the side lanes may fall behind.
# Default: one slot per use in the whole backward for the two-per-block rings (their index runs on across the decoder and
# the encoder) and one per block of the deepest stack for the others: no slot is reused while a lane may still read it
# (1 GB at batch 4 for ViT-B; measured 4.55 -> 4.39 ms per step against rings of 4 / 2 and another 0.03 ms for the
# whole-backward depth).  VITAE_RING / VITAE_RING_BLOCK override (>= 4 / >= 2) when memory matters more.
def ring_depths(enc_depth: int, dec_depth: int):
    ring = int(os.environ.get("VITAE_RING", str(2 * (enc_depth + dec_depth) + 2)))
    ring_block = int(os.environ.get("VITAE_RING_BLOCK", str(max(enc_depth, dec_depth))))
    return max(4, ring), max(2, ring_block)


MAX_INPUT_ADDRESSES = 4   # input-volume addresses that get their own (zero-copy) graphs; others are copied (below)


class _Arena:
    """Named device buffers of one (batch, keep) shape; allocated once, reused every step (CUDA-graph friendly)."""

    def __init__(self, device):
        self.device = device
        self.nbytes = 0

    def new(self, shape, dtype) -> torch.Tensor:
        t = torch.empty(shape, dtype=dtype, device=self.device)
        self.nbytes += t.numel() * t.element_size()
        return t


class _BlockBufs:
    pass


class MAEPlan:
    """Activation arena + static row maps for one (B, keep) shape."""

    def __init__(self, eng: "MAEEngine", B: int, keep: int):
        c = eng.cfg
        dev = eng.device
        a = _Arena(dev)
        self.B, self.keep = B, keep
        L, P = eng.L, eng.P
        self.Ne, self.Nd = keep + 1, L + 1
        self.Me, self.Md = B * self.Ne, B * self.Nd
        self.nmask = L - keep
        self.noise = a.new((B, L), _F32)
        self.ids_shuffle = a.new((B, L), _I32)
        self.ids_restore = a.new((B, L), _I32)
        self.mask = a.new((B, L), _F32)
        self.maps = {"enc_tok_rows": a.new((B * keep,), _I32), "enc_cls_rows": a.new((B,), _I32),
                     "pe_pos_rows": a.new((B * keep,), _I32), "dec_rows_of_enc": a.new((self.Me,), _I32),
                     "dec_pos_rows_of_enc": a.new((self.Me,), _I32),
                     "masked_dec_rows": a.new((max(1, B * self.nmask),), _I32),
                     "masked_pos_rows": a.new((max(1, B * self.nmask),), _I32)}
        self.cols = a.new((B * keep, eng.Kpe), _BF16)
        self.enc = self._stack(a, eng.enc, self.Me, B, self.Ne)
        self.dec = self._stack(a, eng.dec, self.Md, B, self.Nd)
        D, Dd = eng.enc.dim, eng.dec.dim
        self.latent = a.new((self.Me, D), _BF16)
        self.dlatent: Optional[torch.Tensor] = None   # fp32 upstream gradient of the latent (contrastive predictor)
        self.edge_scratch = self.edge_tgt = self.edge_resid = self.edge_out = self.dedge = None
        # fp32 copy of the latent for consumers outside the kernels (contrastive predictor): the cosine loss is sensitive
        # to bf16 rounding of its inputs (3 % gradient error from that rounding alone)
        self.latent32 = a.new((self.Me, D), _F32) if eng.want_latent32 else None
        self.mean_n, self.rstd_n = a.new((self.Me,), _F32), a.new((self.Me,), _F32)
        self.hN = a.new((self.Md, Dd), _BF16)
        self.mean_dn, self.rstd_dn = a.new((self.Md,), _F32), a.new((self.Md,), _F32)
        self.pred = a.new((B, self.Nd, P), _BF16)
        self.pred32: Optional[torch.Tensor] = None   # fp32 copy of pred, allocated when the caller wants fp32 back
        self.step_id = 0
        self.mse_part: Optional[torch.Tensor] = None    # row partials of the loss fused into the decoder_pred GEMM
        self.fused_loss = False                         # the last forward left d recon / d pred (unscaled) in dpred
        self.patch_sums = a.new((B * L,), _F32)
        self.loss_out = a.new((2,), _F32)
        # backward scratch
        Mmax = max(self.Me, self.Md)
        Dmax = max(D, Dd)
        Hmax = max(eng.enc.hidden, eng.dec.hidden)
        self.dloss = a.new((1,), _F32)
        self.dpred = a.new((B, self.Nd, P), _BF16)
        # rings of eng.ring (two uses per block) / eng.ring_block (one use per block) buffers: the side lanes read them
        # (weight-gradient GEMMs, the per-block column reductions) after the main lane has moved on (ring_depths above)
        self.dres = [a.new((Mmax * Dmax,), _F32) for _ in range(eng.ring)]
        self.dres16 = [a.new((Mmax * Dmax,), _BF16) for _ in range(eng.ring)]
        self.d_d = a.new((Mmax * Dmax,), _BF16)
        # LayerNorm-backward inputs: read again by the reduction lane (affine-gradient reductions)
        self.d_ln = [a.new((Mmax * Dmax,), _BF16) for _ in range(eng.ring)]
        bcr = max(ops.block_colreduce_workspace_bytes(m, [st.hidden, 3 * st.dim, st.dim, st.dim, st.dim, st.dim])
                  for m, st in ((self.Me, eng.enc), (self.Md, eng.dec)))
        self.bcr_ws = torch.zeros(bcr, dtype=torch.uint8, device=dev)   # zero-filled: ticket counters (side lane only)
        a.nbytes += bcr
        self.d_hid = [a.new((Mmax * Hmax,), _BF16) for _ in range(eng.ring_block)]
        self.dqkv = [a.new((Mmax * 3 * Dmax,), _BF16) for _ in range(eng.ring_block)]
        self.delta = a.new((B * max(eng.enc.heads * self.Ne, eng.dec.heads * self.Nd),), _F32)
        ln_ws = ops.layernorm_param_grads_workspace_bytes(Mmax, Dmax)
        self.ln_ws = torch.zeros(ln_ws, dtype=torch.uint8, device=dev)   # zero-filled: ticket counters (side lane only)
        a.nbytes += ln_ws
        self.ln_calls = 0
        cs_bytes = max(ops.colsum_workspace_bytes(r, c) for r in (self.Me, self.Md, B * keep)
                       for c in (P, 3 * D, 3 * Dd, eng.enc.hidden, eng.dec.hidden, D, Dd))
        self.colsum_ws = torch.zeros(cs_bytes, dtype=torch.uint8, device=dev)   # zero-filled: ticket counters
        a.nbytes += cs_bytes
        self.g_embed = a.new((self.Me, Dd), _BF16)
        self.g_pe = a.new((B * keep, D), _BF16)
        self.vol_static: Optional[torch.Tensor] = None   # copy target once too many distinct input addresses were seen
        self.graphs: Dict[tuple, _GraphSlot] = {}
        self.nbytes = a.nbytes
        self.vol: Optional[torch.Tensor] = None   # the volume of the current step (read in place by the kernels)

    def edge_buffers(self, eng: "MAEEngine") -> None:
        """Scratch of the fused edge-map loss (allocated on first use: 5 volume-sized fp32 buffers)."""
        if self.edge_scratch is None:
            B, C, V = self.B, eng.C, eng.V
            self.edge_scratch = torch.empty(ops.edge_scratch_floats(B, C, V), dtype=_F32, device=eng.device)
            self.edge_tgt = torch.empty((B, V, V, V), dtype=_F32, device=eng.device)
            self.edge_resid = torch.empty((B, V, V, V), dtype=_F32, device=eng.device)
            self.edge_out = torch.zeros((1,), dtype=_F32, device=eng.device)
            self.dedge = torch.zeros((1,), dtype=_F32, device=eng.device)

    def pred_view(self, dtype) -> torch.Tensor:
        """The reference's ``pred`` [N, L, P] (cls row dropped, vit_autoenc.py:200-201) as a view of the workspace."""
        src = self.pred32 if dtype == _F32 else self.pred
        return src[:, 1:, :]

    @staticmethod
    def _stack(a: _Arena, st: StackSpec, M: int, B: int, N: int):
        D, hid = st.dim, st.hidden
        x = [a.new((M, D), _F32) for _ in range(st.depth + 1)]
        blocks = []
        for _ in range(st.depth):
            b = _BlockBufs()
            b.ln1 = a.new((M, D), _BF16); b.mean1 = a.new((M,), _F32); b.rstd1 = a.new((M,), _F32)
            b.qkv = a.new((M, 3 * D), _BF16); b.o = a.new((M, D), _BF16); b.lse = a.new((B, st.heads, N), _F32)
            b.xmid = a.new((M, D), _F32)
            b.ln2 = a.new((M, D), _BF16); b.mean2 = a.new((M,), _F32); b.rstd2 = a.new((M,), _F32)
            b.pre = a.new((M, hid), _BF16); b.act = a.new((M, hid), _BF16)
            blocks.append(b)
        s = _BlockBufs()
        s.x, s.blocks = x, blocks
        return s


class MAEEngine:
    """Owns flat parameters and per-shape plans of one MaskedAutoencoderViT and runs its forward / backward."""

    def __init__(self, cfg: dict, named_params: Dict[str, torch.nn.Parameter], pos_embed: torch.Tensor,
                 decoder_pos_embed: torch.Tensor, ln_eps: float):
        dev = pos_embed.device
        if dev.type != "cuda":
            raise ops._lib.VitaeError("MAEEngine needs CUDA tensors: this package has no CPU path")
        ops._lib.check(ops._lib.load().vitae_check_device(), "vitae_check_device")
        self.cfg = cfg
        self.device = dev
        V, p, C = cfg["volume_size"], cfg["patch_size"], cfg["in_chans"]
        self.V, self.p, self.C = V, p, C
        self.g = V // p
        self.L = self.g ** 3
        self.P = p ** 3 * C
        self.Kpe = C * p ** 3
        self.eps = float(ln_eps)
        D, Dd = cfg["embed_dim"], cfg["decoder_embed_dim"]
        self.enc = StackSpec("blocks", D, cfg["num_heads"], int(D * cfg["mlp_ratio"]), cfg["depth"])
        self.dec = StackSpec("decoder_blocks", Dd, cfg["decoder_num_heads"], int(Dd * cfg["mlp_ratio"]),
                             cfg["decoder_depth"])
        self.ring, self.ring_block = ring_depths(self.enc.depth, self.dec.depth)
        for st in (self.enc, self.dec):
            if st.head_dim not in (16, 32, 64):
                raise ops._lib.VitaeError(f"unsupported head_dim {st.head_dim} (kernels exist for 16/32/64)")
            if st.dim % 8 or st.hidden % 8 or st.dim > 1024:
                raise ops._lib.VitaeError(f"unsupported width {st.dim}/{st.hidden}")
        if self.P % 8 or p % 4:
            raise ops._lib.VitaeError("patch_size must be a multiple of 4")
        self.has_predictor = "predictor.0.weight" in named_params
        self.flat = FlatParams(named_params, backward_param_order(cfg["depth"], cfg["decoder_depth"], self.has_predictor), dev)
        # BatchNorm1d of the predictor: (running_mean, running_var, eps, momentum), set by ContrastiveMAEViT.engine()
        self.bn_state = None
        self.pos = pos_embed.detach().reshape(self.L + 1, D).contiguous()
        self.dpos = decoder_pos_embed.detach().reshape(self.L + 1, Dd).contiguous()
        self.plans: Dict[Tuple[int, int], MAEPlan] = {}
        self.lanes = _Lanes(dev)
        self.reduce_lane = os.environ.get("VITAE_REDUCE_LANE", "1") != "0"   # column reductions on their own stream
        self.bg_stream: Optional[torch.cuda.Stream] = None     # edge-map target branch underneath the forward
        self._bg_done: Optional[torch.cuda.Event] = None
        self.cap_stream: Optional[torch.cuda.Stream] = None
        self.pf_stream = torch.cuda.Stream(device=dev)
        self.ws_main, self.ws_side = ops.GrowBuf(dev), ops.GrowBuf(dev)
        self.use_graphs = True
        self.use_side_lane = True
        # pull the next block's weights (and, in backward, its saved activations) into the 126 MB L2 while the current
        # block computes: every kernel is a few microseconds long and would otherwise start with a cold HBM load
        # (opt-in, VITAE_L2_PREFETCH=1: measured on B200 at batch 4 it does not pay -- 4.71 ms/step with, 4.65 without)
        self.use_l2_prefetch = os.environ.get("VITAE_L2_PREFETCH", "0") == "1"
        self.graph_replayed_launches = 0   # kernels executed through graph replays (vitae_launch_count sees enqueues)
        self.optim: Optional["FusedAdamW"] = None
        self.want_latent32 = False         # set before the first plan is built (ContrastiveMAEViT)
        self.edge_taps = ops.gaussian_taps(2.0)   # sigma = 2 at the call site, model/vit_autoenc.py:222
        # Parameter groups in FORWARD order (contiguous slices of the flat buffers, which are laid out in backward order):
        # the optimizer can update them one after the other on its own stream while the next forward, which waits for
        # group g right before its first kernel that reads it, is already running (FusedAdamW.step(overlap=True)).
        # The events are "external": inside a captured forward they become event-wait nodes on the latest record.
        self.group_of_param, self.group_ranges = self._make_param_groups()
        self.param_ready = [torch.cuda.Event(external=True) for _ in self.group_ranges]
        self.opt_stream = torch.cuda.Stream(device=dev, priority=0)
        self.params_in_flight = False      # an overlapped optimizer step may still be writing the parameters
        self._waited = set()               # groups the pass being enqueued has already waited for
        self.grad_buckets = None
        self.bucket_elems = 32 << 20       # 128 MB of fp32 gradients per all-reduce bucket
        # data parallel, sharded step (dp.ShardedStep): utils.misc.NativeScalerWithGradNormCount announces before the
        # backward whether its fused optimizer step follows ("reduce": the backward reduces each finished slice into its
        # owner) or not ("skip": gradient accumulation, nothing is exchanged yet); None = all-reduce inside backward
        self.supports_sharded = True
        self.defer_exchange: Optional[str] = None
        self.grads_local = False           # the flat gradient buffer holds rank-local sums that still need the exchange
        # squared-norm partials of the gradient slices, computed per backward part on a stream of their own underneath the
        # later parts (FusedAdamW.step then starts with the tiny finalize instead of an 80 us pass over all gradients)
        self.norm_stream = torch.cuda.Stream(device=dev, priority=0)
        self.norm_blocks = 148
        self.norm_ws = torch.zeros(16 * self.norm_blocks, dtype=_F32, device=dev)
        self.norm_partials = 0             # > 0: norm_ws[:norm_partials] covers the whole gradient buffer as it is now
        self.g16: Optional[torch.Tensor] = None            # bf16 staging of the gradient exchange (allocated for N > 1)
        self.comm_stream: Optional[torch.cuda.Stream] = None

    # ------------------------------------------------------------------------------------------------ helpers
    def _make_param_groups(self):
        """name -> group id, and per group the [start, end) element range of the flat buffers (64-aligned)."""
        eb = max(1, math.ceil(self.enc.depth / 4))
        db = max(1, math.ceil(self.dec.depth / 2))
        n_enc = math.ceil(self.enc.depth / eb) if self.enc.depth else 0
        n_dec = math.ceil(self.dec.depth / db) if self.dec.depth else 0

        def gid(name: str) -> int:
            if name.startswith("patch_embed.") or name == "cls_token":
                return 0
            if name.startswith("blocks."):
                return 1 + int(name.split(".")[1]) // eb
            if name in ("norm.weight", "norm.bias", "decoder_embed.weight", "decoder_embed.bias", "mask_token"):
                return 1 + n_enc
            if name.startswith("decoder_blocks."):
                return 2 + n_enc + int(name.split(".")[1]) // db
            return 2 + n_enc + n_dec          # decoder_norm, decoder_pred (and the contrastive predictor in front of them)
        group_of = {n: gid(n) for n in self.flat.order}
        ngroups = 3 + n_enc + n_dec
        ranges = [[None, None] for _ in range(ngroups)]
        names = self.flat.order
        for i, n in enumerate(names):
            o, k, _ = self.flat.offsets[n]
            end = self.flat.offsets[names[i + 1]][0] if i + 1 < len(names) else self.flat.total
            r = ranges[group_of[n]]
            r[0] = o if r[0] is None else min(r[0], o)
            r[1] = end if r[1] is None else max(r[1], end)
        spans = sorted((r[0], r[1]) for r in ranges)
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and spans[0][0] == 0 and spans[-1][1] == self.flat.total, \
            "parameter groups must tile the flat buffer"
        return group_of, [tuple(r) for r in ranges]

    def _block_p16(self, prefix: str, i: int) -> torch.Tensor:
        """bf16 shadow of block i's parameters: one contiguous slice of the flat buffer (fc2.weight ... norm1.bias)."""
        a = self.flat.offsets[f"{prefix}.{i}.mlp.fc2.weight"][0]
        o, k, _ = self.flat.offsets[f"{prefix}.{i}.norm1.bias"]
        return self.flat.p16[a:o + k]

    def _prefetch(self, tensors) -> None:
        """L2 prefetch on its own stream (forked from the current point of the main lane, joined by _Lanes.join): it must
        not queue behind the side lane's weight-gradient GEMMs."""
        if self.use_l2_prefetch and self.use_side_lane:
            ts = [t for t in tensors if t is not None]
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(main)
            self.pf_stream.wait_event(ev)
            with torch.cuda.stream(self.pf_stream):
                ops.prefetch_l2(ts)
            self.lanes.extra_dirty = self.pf_stream

    def _prefetch_block(self, st: StackSpec, sb, i: int, with_acts: bool) -> None:
        if 0 <= i < st.depth:
            ts = [self._block_p16(st.prefix, i)]
            if with_acts:
                b = sb.blocks[i]
                ts += [b.pre, b.act, b.ln2, b.o, b.qkv, b.ln1, b.xmid, sb.x[i]]
            self._prefetch(ts)

    def _capture_stream(self) -> Optional[torch.cuda.Stream]:
        """Graphs are captured on a stream of priority VITAE_MAIN_PRIORITY (kernel nodes inherit it): above the side lanes
        and the background stream (0, the lowest)."""
        prio = int(os.environ.get("VITAE_MAIN_PRIORITY", "-1"))
        if prio == 0:
            return None
        if self.cap_stream is None:
            self.cap_stream = torch.cuda.Stream(device=self.device, priority=prio)
        return self.cap_stream

    def _background(self, fn) -> None:
        """Forks ``fn`` onto the background stream (priority 0, below the graph-captured main lane: it fills SMs the main
        chain leaves idle); ``_background_join`` orders the main lane after it.  VITAE_EDGE_OVERLAP=0 runs it inline."""
        if os.environ.get("VITAE_EDGE_OVERLAP", "1") == "0":
            fn()
            return
        if self.bg_stream is None:
            self.bg_stream = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.bg_stream.wait_event(ev)
        with torch.cuda.stream(self.bg_stream):
            fn()
        self._bg_done = torch.cuda.Event()
        self._bg_done.record(self.bg_stream)

    def _background_join(self) -> None:
        if self._bg_done is not None:
            torch.cuda.current_stream().wait_event(self._bg_done)
            self._bg_done = None

    def _need(self, name: str) -> None:
        """Orders the current stream after the optimizer's update of the group that holds parameter ``name`` (once per
        group and pass: an event-wait node in front of a kernel costs that kernel its programmatic launch edge)."""
        g = self.group_of_param[name]
        if g not in self._waited:
            self._waited.add(g)
            torch.cuda.current_stream().wait_event(self.param_ready[g])

    def wait_params(self) -> None:
        """Current stream waits for every parameter group (call before non-engine work touches parameters, gradients or
        optimizer state after an overlapped optimizer step)."""
        if self.params_in_flight:
            for ev in self.param_ready:
                torch.cuda.current_stream().wait_event(ev)
            self.params_in_flight = False

    def plan(self, B: int, keep: int, slot: int = 0) -> MAEPlan:
        """Activation arena for this shape; ``slot`` > 0 gives a second, independent arena (the contrastive model keeps the
        encoder activations of both views alive until backward)."""
        key = (B, keep) if slot == 0 else (B, keep, slot)
        pl = self.plans.get(key)
        if pl is None:
            pl = MAEPlan(self, B, keep)
            self.plans[key] = pl
        return pl

    def _w(self, name):   # bf16 GEMM operand
        return self.flat.v16[name]

    def _p(self, name):   # fp32 master (bias / LayerNorm affine / tokens)
        return self.flat.v32[name]

    def _g(self, name):   # fp32 gradient
        return self.flat.vg[name]

    def _side(self, fn, reads=(), lane: int = 0):
        if self.use_side_lane:
            self.lanes.side(fn, reads, lane if self.reduce_lane else 0)
        else:
            fn()

    def _run(self, pl: MAEPlan, key: tuple, fn) -> None:
        """Runs ``fn`` (a fixed sequence of kernel enqueues over static buffers): eagerly the first time a key is seen,
        captured into a CUDA graph the second time, replayed afterwards."""
        if not self.use_graphs or torch.cuda.is_current_stream_capturing():
            fn()
            return
        slot = pl.graphs.get(key)
        if slot is None:
            slot = pl.graphs[key] = _GraphSlot()
        slot.calls += 1
        if slot.calls == 1:
            fn()
            return
        if slot.graph is None:
            lib = ops._lib.load()
            n0 = lib.vitae_launch_count()
            g = torch.cuda.CUDAGraph()
            # no cyclic garbage collection while capturing: destructors with CUDA side effects (graphs, symmetric-memory
            # mappings of a model that died earlier) invalidate the capture; torch.cuda.graph collects once on entry
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g, stream=self._capture_stream()):
                    fn()
            finally:
                if gc_was_on:
                    gc.enable()
            slot.launches = lib.vitae_launch_count() - n0
            self.graph_replayed_launches -= slot.launches    # the capture pass enqueued them without executing
            slot.graph = g
        slot.graph.replay()
        self.graph_replayed_launches += slot.launches

    def _resident(self, pl: MAEPlan, vol: torch.Tensor) -> torch.Tensor:
        """The kernels read the caller's volume in place.  CUDA graphs bake its address in: the first address a plan
        sees is used in place (a resident batch trains with zero copies); volumes at any other address are copied
        into one static buffer (a device-to-device copy is ~40 us for 4 x 128^3 x 4 fp32, a graph capture ~100 ms)."""
        if not self.use_graphs:
            return vol
        ptr = vol.data_ptr()
        known = {k[1] for k in pl.graphs if k[0] in ("fwd", "enc")}
        if ptr in known or len(known) < MAX_INPUT_ADDRESSES:
            return vol
        if pl.vol_static is None:
            pl.vol_static = torch.empty_like(vol)
        if ptr != pl.vol_static.data_ptr():
            pl.vol_static.copy_(vol)
        return pl.vol_static

    # ------------------------------------------------------------------------------------------------ forward
    def forward_encoder_only(self, vol: torch.Tensor, noise: torch.Tensor, keep: int, slot: int = 1,
                             with_predictor: bool = False) -> MAEPlan:
        """Encoder pass alone (second view of the contrastive model, model/vit_autoenc.py:277) into arena ``slot``."""
        pl = self.plan(vol.shape[0], keep, slot)
        vol = self._resident(pl, vol)
        pl.vol = vol
        pl.noise.copy_(noise)
        self.refresh_shadow()
        if with_predictor:
            self._pred_bufs(pl)
        def body():
            self.encode(pl, vol, pl.noise)
            if with_predictor:
                self.predictor_forward(pl)
            self.lanes.join()
        self._run(pl, ("enc", vol.data_ptr(), with_predictor), body)
        return pl

    def forward(self, vol: torch.Tensor, noise: torch.Tensor, keep: int, want_loss: bool = True,
                pred_f32: bool = False, want_edge: bool = False, with_predictor: bool = False) -> MAEPlan:
        """vol fp32 [B,C,V,V,V] (contiguous, CUDA); noise fp32 [B,L].  Fills plan.pred / mask / loss_out (and, with
        ``want_edge``, plan.edge_out = raw edge-map loss of model/vit_autoenc.py:221-224)."""
        B = vol.shape[0]
        pl = self.plan(B, keep)
        vol = self._resident(pl, vol)
        pl.vol = vol
        pl.noise.copy_(noise)
        if pred_f32 and pl.pred32 is None:
            pl.pred32 = torch.empty((pl.B, pl.Nd, self.P), dtype=_F32, device=self.device)
        self.refresh_shadow()
        if want_edge:
            pl.edge_buffers(self)
        if with_predictor:
            self._pred_bufs(pl)
        # the reconstruction loss rides in the epilogue of the decoder_pred GEMM (vitae_gemm_pred_mse) whenever nothing needs
        # the separate kernels: 4 channels, patch % 8 == 0, no fp32 copy of pred, no edge-map term reading pred
        fuse_loss = (want_loss and not want_edge and not pred_f32 and self.C == 4 and self.p % 8 == 0 and self.P % 128 == 0
                     and pl.nmask > 0 and os.environ.get("VITAE_FUSED_LOSS", "1") != "0")
        if fuse_loss and pl.mse_part is None:
            pl.mse_part = torch.empty(ops.pred_mse_partial_floats(pl.Md, self.P, 128), dtype=_F32, device=self.device)
        pl.fused_loss = fuse_loss

        def body():
            if want_edge:
                # the target branch (blur + Sobel of the input volume) depends on nothing the model computes: it runs on
                # the background stream underneath the encoder / decoder, whose small GEMMs leave most SMs idle
                self._background(lambda: ops.edge_target(vol, self.edge_taps, pl.edge_scratch, pl.edge_tgt))
            self.encode(pl, vol, pl.noise)
            if with_predictor:
                self.predictor_forward(pl)
            self.decode(pl, pred_f32, loss_vol=vol if fuse_loss else None)
            if want_loss and not fuse_loss:
                ops.masked_mse_fwd(pl.pred, vol, pl.mask, pl.patch_sums, pl.loss_out, self.p)
            if want_edge:
                self._background_join()
                ops.edge_loss_fwd(pl.pred, pl.edge_tgt, pl.edge_scratch, pl.edge_resid, pl.edge_out, pl.B, self.C, self.V,
                                  self.p)
            self.lanes.join()          # the side lane carries the L2 prefetches of the forward
        self._run(pl, ("fwd", vol.data_ptr(), pred_f32, want_loss, want_edge, with_predictor, fuse_loss), body)
        self.params_in_flight = False      # the pass above waited for every parameter group
        return pl

    def refresh_shadow(self) -> None:
        """bf16 shadow of the parameters; a cast pass (parameters written through torch) reads all of them."""
        if self.flat.shadow_version != self.flat._version():
            self.wait_params()
        self.flat.refresh_shadow()

    def encode(self, pl: MAEPlan, vol: torch.Tensor, noise: torch.Tensor) -> None:
        """model/vit_autoenc.py:157-177 (forward_encoder): patch embed of the kept patches only (the dropped ones are
        discarded by the gather at :147, so embedding them is dead work) + pos, cls row, blocks, final norm."""
        B, keep, D = pl.B, pl.keep, self.enc.dim
        self._waited = set()
        self._prefetch([self._w("patch_embed.proj.weight")] + ([self._block_p16("blocks", 0)] if self.enc.depth else []))
        ops.random_masking(noise, pl.ids_shuffle, pl.ids_restore, pl.mask, keep)
        ops.build_row_maps(pl.ids_shuffle, keep, pl.maps)
        ops.im2col_patches(vol, pl.ids_shuffle, pl.cols, self.p, keep)
        x0 = pl.enc.x[0]
        self._need("patch_embed.proj.weight")
        ops.gemm(pl.cols, self._w("patch_embed.proj.weight"), B * keep, D, self.Kpe,
                 bias=self._p("patch_embed.proj.bias"), addend=self.pos, add_rows=pl.maps["pe_pos_rows"], ldadd=D,
                 out_f32=x0, out_rows=pl.maps["enc_tok_rows"], workspace=self.ws_main)
        ops.fill_rows(x0, pl.maps["enc_cls_rows"], B, D, self._p("cls_token"), None, self.pos, None)
        self._stack_fwd(self.enc, pl.enc, pl.Me, B, pl.Ne,
                        after=[self._w("decoder_embed.weight")] + ([self._block_p16("decoder_blocks", 0)] if self.dec.depth else []))
        self._need("norm.weight")
        ops.layernorm_fwd(pl.enc.x[-1], self._p("norm.weight"), self._p("norm.bias"), pl.latent, pl.mean_n, pl.rstd_n,
                          self.eps, y_f32=pl.latent32)

    def decode(self, pl: MAEPlan, pred_f32: bool = False, loss_vol: Optional[torch.Tensor] = None) -> None:
        """model/vit_autoenc.py:179-203 (forward_decoder); pl.pred keeps the cls row (row 0 of each sample).  ``loss_vol``:
        evaluate the masked reconstruction loss against this volume in the decoder_pred epilogue (fills pl.loss_out and
        pl.dpred = d recon / d pred before the upstream factor)."""
        if pred_f32 and pl.pred32 is None:
            pl.pred32 = torch.empty((pl.B, pl.Nd, self.P), dtype=_F32, device=self.device)
        B, D, Dd = pl.B, self.enc.dim, self.dec.dim
        xd0 = pl.dec.x[0]
        self._waited.discard(self.group_of_param["decoder_embed.weight"])   # decode() may be called on its own
        self._need("decoder_embed.weight")
        ops.gemm(pl.latent, self._w("decoder_embed.weight"), pl.Me, Dd, D, bias=self._p("decoder_embed.bias"),
                 addend=self.dpos, add_rows=pl.maps["dec_pos_rows_of_enc"], ldadd=Dd, out_f32=xd0,
                 out_rows=pl.maps["dec_rows_of_enc"], workspace=self.ws_main)
        if pl.nmask > 0:
            ops.fill_rows(xd0, pl.maps["masked_dec_rows"], B * pl.nmask, Dd, self._p("mask_token"), None, self.dpos,
                          pl.maps["masked_pos_rows"])
        self._stack_fwd(self.dec, pl.dec, pl.Md, B, pl.Nd, after=[self._w("decoder_pred.weight")])
        self._need("decoder_norm.weight")
        ops.layernorm_fwd(pl.dec.x[-1], self._p("decoder_norm.weight"), self._p("decoder_norm.bias"), pl.hN,
                          pl.mean_dn, pl.rstd_dn, self.eps)
        if loss_vol is not None:
            mask_sum = float(pl.B * pl.nmask)          # every sample removes exactly L - keep patches (vit_autoenc.py:137-153)
            self.lanes.before_write("dpred")
            ops.gemm_pred_mse(pl.hN, self._w("decoder_pred.weight"), self._p("decoder_pred.bias"), B, self.L, Dd, loss_vol,
                              pl.mask, self.p, mask_sum, pl.pred.view(pl.Md, self.P), pl.dpred.view(pl.Md, self.P), pl.mse_part)
            # the loss value itself is off the dependency chain: reduce the row partials on a side lane
            self._side(lambda: ops.pred_mse_finalize(pl.mse_part, self.P, mask_sum, pl.loss_out), lane=1)
            return
        ops.gemm(pl.hN, self._w("decoder_pred.weight"), pl.Md, self.P, Dd, bias=self._p("decoder_pred.bias"),
                 out_bf16=pl.pred.view(pl.Md, self.P), out_f32=pl.pred32.view(pl.Md, self.P) if pred_f32 else None,
                 workspace=self.ws_main)

    def _stack_fwd(self, st: StackSpec, sb, M: int, B: int, N: int, after=()) -> None:
        """model/vit.py:139-144 (Block), :112-124 (Attention), :90-96 (Mlp3D).  ``after``: weights used right after the
        stack (prefetched into L2 during the last block)."""
        D, hid, H, hd = st.dim, st.hidden, st.heads, st.head_dim
        scale = hd ** -0.5
        ws = self.ws_main
        for i in range(st.depth):
            pre = f"{st.prefix}.{i}"
            b, x_in, x_out = sb.blocks[i], sb.x[i], sb.x[i + 1]
            if i + 1 < st.depth:
                self._prefetch_block(st, sb, i + 1, with_acts=False)
            elif after:
                self._prefetch(list(after))
            self._need(f"{pre}.norm1.weight")
            ops.layernorm_fwd(x_in, self._p(f"{pre}.norm1.weight"), self._p(f"{pre}.norm1.bias"), b.ln1, b.mean1,
                              b.rstd1, self.eps)
            ops.gemm(b.ln1, self._w(f"{pre}.attn.qkv.weight"), M, 3 * D, D, bias=self._p(f"{pre}.attn.qkv.bias"),
                     out_bf16=b.qkv, workspace=ws)
            ops.attention_fwd(b.qkv, b.o, b.lse, B, N, H, hd, scale)
            ops.gemm(b.o, self._w(f"{pre}.attn.proj.weight"), M, D, D, bias=self._p(f"{pre}.attn.proj.bias"),
                     addend=x_in, out_f32=b.xmid, workspace=ws)
            ops.layernorm_fwd(b.xmid, self._p(f"{pre}.norm2.weight"), self._p(f"{pre}.norm2.bias"), b.ln2, b.mean2,
                              b.rstd2, self.eps)
            ops.gemm(b.ln2, self._w(f"{pre}.mlp.fc1.weight"), M, hid, D, bias=self._p(f"{pre}.mlp.fc1.bias"),
                     out_bf16=b.pre, out_gelu_bf16=b.act, workspace=ws)
            ops.gemm(b.act, self._w(f"{pre}.mlp.fc2.weight"), M, D, hid, bias=self._p(f"{pre}.mlp.fc2.bias"),
                     addend=b.xmid, out_f32=x_out, workspace=ws)

    # ------------------------------------------------------------------------------------------------ backward
    def backward(self, pl: MAEPlan, dloss: Optional[torch.Tensor], dpred_extra: Optional[torch.Tensor] = None,
                 accumulate: bool = False, sync_grads: bool = False, dlatent: Optional[torch.Tensor] = None,
                 encoder_only: bool = False, dedge: Optional[torch.Tensor] = None) -> None:
        """Gradient of (dloss * recon_loss [+ <dpred_extra, pred>]) w.r.t. every trainable parameter, written to
        (accumulate=False) or added into (True) the flat gradient buffer.  Hand-derived reverse of forward().

        ``sync_grads`` (data parallel, world size > 1): the backward runs as a few stages (one CUDA graph each) whose
        gradients are contiguous slices of the flat buffer; the mean all-reduce of a stage's slice is issued as soon as
        that stage has been enqueued and overlaps the later stages (NCCL runs on its own stream).

        ``dlatent`` fp32 [B*Ne, D]: an additional upstream gradient of the normalised encoder output (the contrastive
        predictor consumes it, model/vit_autoenc.py:280-283).  ``encoder_only``: reverse of forward_encoder_only() -- the
        only upstream gradient is ``dlatent``; decoder parameters are not touched."""
        self.wait_params()         # an overlapped optimizer step of parameter groups this step's forward did not touch
        if dloss is None:
            pl.dloss.zero_()
        else:
            pl.dloss.copy_(dloss.reshape(1))
        if dedge is not None:      # upstream gradient of the raw edge-map loss (forward(want_edge=True) must have run)
            if pl.edge_resid is None:
                raise ops._lib.VitaeError("edge-map gradient without an edge-map forward")
            pl.dedge.copy_(dedge.reshape(1))
        if dlatent is not None:
            if pl.dlatent is None:
                pl.dlatent = torch.empty((pl.Me, self.enc.dim), dtype=_F32, device=self.device)
            pl.dlatent.copy_(dlatent.reshape(pl.Me, self.enc.dim))
        elif encoder_only:
            raise ops._lib.VitaeError("encoder-only backward needs the latent gradient")
        staged = sync_grads and dp.world_size() > 1
        stages = self._backward_stages(pl, dpred_extra, accumulate, split=staged, with_dlatent=dlatent is not None,
                                       encoder_only=encoder_only, with_edge=dedge is not None)
        reducer = None
        if staged:
            if self.defer_exchange == "reduce" and self.flat.sharded is not None:
                reducer = self.flat.sharded      # owner-side reduce of each slice instead of an all-reduce
                reducer.begin()
            else:
                reducer = self._grad_reducer()
        self._norm_count = 0
        self.norm_partials = 0
        for i, (fn, (a, b)) in enumerate(stages):
            if dpred_extra is not None and i == 0:   # auxiliary torch-side terms that consume ``pred``: not graphed
                fn()
            else:
                self._run(pl, ("bwd", i, len(stages), pl.vol.data_ptr(), bool(accumulate), dlatent is not None,
                               encoder_only, dedge is not None, pl.fused_loss), fn)
            if reducer is not None:
                if reducer is self.flat.sharded:
                    reducer.launch(self.flat.g32[a:b], final=i == len(stages) - 1)
                else:
                    reducer.launch(self.flat.g32[a:b])
        if reducer is not None:
            reducer.wait()
        if len(stages) == 1 and not staged and not encoder_only and dpred_extra is None:
            # (a replayed graph does not run the python above: the count is kept from the pass that was captured)
            key = ("bwd", 0, 1, pl.vol.data_ptr(), bool(accumulate), dlatent is not None, encoder_only, dedge is not None,
                   pl.fused_loss)
            if self._norm_count:
                pl.norm_counts = getattr(pl, "norm_counts", {})
                pl.norm_counts[key] = self._norm_count
            self.norm_partials = getattr(pl, "norm_counts", {}).get(key, 0)

    def _backward_stages(self, pl: MAEPlan, dpred_extra: Optional[torch.Tensor], acc: bool, split: bool,
                         with_dlatent: bool = False, encoder_only: bool = False, with_edge: bool = False):
        """[(enqueue function, (start, end) slice of the flat gradient buffer it completes)], in execution order.
        split=False: one stage.  The stages share ``state['cur']`` (which of the two residual-gradient buffers is live)."""
        B, D, Dd, P = pl.B, self.enc.dim, self.dec.dim, self.P
        cws, wsm, wss = pl.colsum_ws, self.ws_main, self.ws_side
        lanes = self.lanes
        state = {"cur": 0}
        # encoder blocks in groups of (about) 3, top first: the first group shares a stage with the decoder embed / encoder
        # norm, every further group is a stage of its own (finer slices near the end of backward leave less of the
        # gradient exchange exposed after the last kernel)
        gsz = max(1, int(os.environ.get("VITAE_DP_GROUP", "3")))
        enc_desc = list(reversed(range(self.enc.depth)))
        enc_groups = [enc_desc[k:k + gsz] for k in range(0, len(enc_desc), gsz)] or [[]]
        enc_hi, enc_rest = enc_groups[0], enc_groups[1:]
        dec_all = list(reversed(range(self.dec.depth)))

        # the forward's fused loss epilogue already left g = d recon / d pred in dpred, without the upstream factor: the
        # GEMMs below apply it through alpha_ptr.  Any other term that writes into dpred needs the scaled gradient instead.
        use_g = pl.fused_loss and dpred_extra is None and not with_edge
        up = pl.dloss if use_g else None

        def stage_pred():
            self._prefetch_block(self.dec, pl.dec, self.dec.depth - 1, with_acts=True)
            if not use_g:
                # ---- loss: d recon / d pred (model/vit_autoenc.py:226-227), zeros for kept patches and the cls row
                lanes.before_write("dpred")
                ops.masked_mse_bwd(pl.pred, pl.vol, pl.mask, pl.loss_out[1:], pl.dloss, pl.dpred, self.p)
                if with_edge:       # + dedge * d raw_edge / d pred (transposed Sobel stencil over the kept residual)
                    ops.edge_loss_bwd(pl.edge_resid, pl.edge_scratch, pl.dedge, pl.dpred, pl.B, self.C, self.V, self.p)
                if dpred_extra is not None:
                    pl.dpred[:, 1:, :].add_(dpred_extra.to(_BF16))
            dpred = pl.dpred.view(pl.Md, P)
            # ---- decoder_pred (vit_autoenc.py:198)

            def side_pred():
                ops.gemm(dpred, pl.hN, P, Dd, pl.Md, a_mn_major=True, b_mn_major=True,
                         out_f32=self._g("decoder_pred.weight"), accumulate=acc, alpha_ptr=up, workspace=wss)
                ops.colsum(dpred, pl.Md, P, self._g("decoder_pred.bias"), cws, accumulate=acc, scale_ptr=up)
            self._side(side_pred, reads=("dpred",))
            d_d = self._ln_in(pl, pl.Md, Dd)
            ops.gemm(dpred, self._w("decoder_pred.weight"), pl.Md, Dd, P, b_mn_major=True, out_bf16=d_d, alpha_ptr=up,
                     workspace=wsm)
            last_dec = f"decoder_blocks.{self.dec.depth - 1}.mlp.fc2.bias" if self.dec.depth else None
            state["cur"] = self._ln_bwd(pl, d_d, pl.dec.x[-1], "decoder_norm", pl.mean_dn, pl.rstd_dn, None, 0, pl.Md, Dd,
                                        acc, last_dec)

        def stage_dec():
            state["cur"] = self._stack_bwd(self.dec, pl.dec, pl, pl.Md, B, pl.Nd, state["cur"], acc, dec_all)

        def stage_mid():
            cur = state["cur"]
            self._prefetch([self._w("decoder_embed.weight"), pl.latent])
            self._prefetch_block(self.enc, pl.enc, self.enc.depth - 1, with_acts=True)
            dxd = pl.dres[cur][:pl.Md * Dd].view(pl.Md, Dd)
            # ---- mask tokens, decoder_embed (vit_autoenc.py:181-190)
            lanes.before_write("g_embed")
            ops.gather_rows(dxd, pl.maps["dec_rows_of_enc"], pl.Me, Dd, pl.g_embed, None)

            def side_embed():
                if pl.nmask > 0:
                    ops.sum_rows(dxd, pl.maps["masked_dec_rows"], B * pl.nmask, Dd, self._g("mask_token").view(-1), acc)
                elif not acc:
                    self._g("mask_token").zero_()
                ops.gemm(pl.g_embed, pl.latent, Dd, D, pl.Me, a_mn_major=True, b_mn_major=True,
                         out_f32=self._g("decoder_embed.weight"), accumulate=acc, workspace=wss)
                ops.colsum(pl.g_embed, pl.Me, Dd, self._g("decoder_embed.bias"), cws, accumulate=acc)
            self._side(side_embed, reads=("g_embed", ("dres", cur)))
            d_e = self._ln_in(pl, pl.Me, D)
            ops.gemm(pl.g_embed, self._w("decoder_embed.weight"), pl.Me, D, Dd, b_mn_major=True, out_bf16=d_e,
                     workspace=wsm)
            # ---- encoder norm + the upper half of the blocks (vit_autoenc.py:172-175)
            last_enc = f"blocks.{self.enc.depth - 1}.mlp.fc2.bias" if self.enc.depth else None
            cur = self._ln_bwd(pl, d_e, pl.enc.x[-1], "norm", pl.mean_n, pl.rstd_n, None, 0, pl.Me, D, acc, last_enc,
                               dy2=pl.dlatent if with_dlatent else None)
            state["cur"] = self._stack_bwd(self.enc, pl.enc, pl, pl.Me, B, pl.Ne, cur, acc, enc_hi)

        def stage_enc_top():   # encoder-only pass: the latent gradient is the only upstream gradient of the final norm
            self._prefetch_block(self.enc, pl.enc, self.enc.depth - 1, with_acts=True)
            last_enc = f"blocks.{self.enc.depth - 1}.mlp.fc2.bias" if self.enc.depth else None
            cur = self._ln_bwd(pl, pl.dlatent, pl.enc.x[-1], "norm", pl.mean_n, pl.rstd_n, None, 0, pl.Me, D, acc, last_enc)
            state["cur"] = self._stack_bwd(self.enc, pl.enc, pl, pl.Me, B, pl.Ne, cur, acc, enc_hi)

        def stage_enc_group(layers):
            def run():
                state["cur"] = self._stack_bwd(self.enc, pl.enc, pl, pl.Me, B, pl.Ne, state["cur"], acc, layers)
            return run

        def stage_embed():
            dx0 = pl.dres[state["cur"]][:pl.Me * D].view(pl.Me, D)
            # ---- cls token, patch embed (vit_autoenc.py:160-170); no input gradient for the volume
            lanes.before_write("g_pe")
            ops.gather_rows(dx0, pl.maps["enc_tok_rows"], B * pl.keep, D, pl.g_pe, None)
            ops.sum_rows(dx0, pl.maps["enc_cls_rows"], B, D, self._g("cls_token").view(-1), acc)
            ops.gemm(pl.g_pe, pl.cols, D, self.Kpe, B * pl.keep, a_mn_major=True, b_mn_major=True,
                     out_f32=self._g("patch_embed.proj.weight").view(D, self.Kpe), accumulate=acc, workspace=wsm)

            def side_pe():
                ops.colsum(pl.g_pe, B * pl.keep, D, self._g("patch_embed.proj.bias"), cws, accumulate=acc)
            self._side(side_pe, reads=("g_pe",))

        off = lambda n: self.flat.offsets[n][0]
        if encoder_only:
            parts = [(stage_enc_top, 0)]       # slices only matter for the gradient exchange; see backward()
        else:
            parts = [(stage_pred, 0)]          # from the start of the buffer: the predictor's gradients ride with this slice
            if dec_all:
                parts.append((stage_dec, off(f"decoder_blocks.{dec_all[0]}.mlp.fc2.weight")))
            parts.append((stage_mid, off("mask_token")))
        for grp in enc_rest:
            parts.append((stage_enc_group(grp), off(f"blocks.{grp[0]}.mlp.fc2.weight")))
        parts.append((stage_embed, off("cls_token")))

        starts = [o for _, o in parts] + [self.flat.total]
        assert starts[0] == 0 and all(a < b for a, b in zip(starts, starts[1:])), "stage slices must tile the gradient buffer"
        norm_parts = (self.use_side_lane and not split and not encoder_only and len(parts) <= 16
                      and os.environ.get("VITAE_NORM_PARTS", "0") == "1")   # opt-in: measured time-neutral (profiles/r02o_ab.txt)

        def joined(fns, with_norm=False):
            def run():
                off_p = 0
                for k, f in enumerate(fns):
                    f()
                    if with_norm:     # this part's slice of the gradient buffer is final once its lanes have drained
                        a, b = starts[k], starts[k + 1]
                        nb = ops.grad_sqnorm_blocks(b - a, self.norm_blocks)
                        lanes.after_all(self.norm_stream, lambda a=a, b=b, o=off_p: ops.grad_sqnorm(
                            self.flat.g32[a:b], self.norm_ws[o:], self.norm_blocks))
                        off_p += nb
                lanes.join()      # a stage is self-contained: its side-lane work is part of it
                if with_norm:
                    self._norm_count = off_p
            return run
        if not split:
            return [(joined([f for f, _ in parts], with_norm=norm_parts), (0, self.flat.total))]
        return [(joined([f]), (starts[i], starts[i + 1])) for i, (f, _) in enumerate(parts)]

    def _ln_in(self, pl: MAEPlan, M: int, D: int) -> torch.Tensor:
        """Buffer for the input gradient (dy) of the NEXT _ln_bwd call; the side lane reads it after the main lane has
        moved on, so main waits here only for the side reader of two calls ago."""
        k = pl.ln_calls % self.ring
        self.lanes.before_write(("d_ln", k))
        return pl.d_ln[k][:M * D].view(M, D)

    def _ln_bwd(self, pl: MAEPlan, dy: torch.Tensor, x: torch.Tensor, name: str, mean, rstd, dx_in_idx: Optional[int],
                out_idx: int, M: int, D: int, acc: bool, bias_name: Optional[str], dy2: Optional[torch.Tensor] = None,
                jobs: Optional[list] = None, reads: Optional[list] = None) -> int:
        """LayerNorm backward.  Main lane (critical path): dres[out_idx] = (dres[dx_in_idx] if given) + LN'(dy), plus its
        bf16 copy dres16[out_idx].  Side lane: the column reductions -- affine gradients and ``bias_name`` (the bias whose
        gradient is the column sum of the new residual gradient).  ``dy`` must come from _ln_in().  Returns out_idx."""
        k = pl.ln_calls % self.ring
        pl.ln_calls += 1
        dx_in = None if dx_in_idx is None else pl.dres[dx_in_idx][:M * D].view(M, D)
        dx_out = pl.dres[out_idx][:M * D].view(M, D)
        dx16 = pl.dres16[out_idx][:M * D].view(M, D)
        self.lanes.before_write(("dres", out_idx), ("dres16", out_idx))
        ops.layernorm_bwd(dy, x, self._p(f"{name}.weight"), mean, rstd, dx_in, dx_out, dx16, dy2=dy2)
        gb = self._g(bias_name) if bias_name is not None else None
        if jobs is not None:       # deferred: the caller launches one column-reduction kernel for the whole block
            jobs.append(ops.col_job(dy, self._g(f"{name}.weight"), x=x, mean=mean, rstd=rstd, a2=dy2,
                                    out1=self._g(f"{name}.bias"), cols=D))
            if gb is not None:
                jobs.append(ops.col_job(dx_out, gb, cols=D))
            reads.extend([("d_ln", k), ("dres", out_idx)])
            return out_idx

        def side():
            ops.layernorm_param_grads(dy, x, mean, rstd, dx_out if gb is not None else None, pl.ln_ws,
                                      dgamma=self._g(f"{name}.weight"), dbeta=self._g(f"{name}.bias"), dbias=gb,
                                      accumulate=acc, dy2=dy2)
        self._side(side, reads=(("d_ln", k), ("dres", out_idx)))
        return out_idx

    def _stack_bwd(self, st: StackSpec, sb, pl: MAEPlan, M: int, B: int, N: int, cur: int, acc: bool, layers) -> int:
        """Reverse of _stack_fwd.  On entry dres[cur] / dres16[cur] hold the gradient w.r.t. the stack output.
        Main lane: dgrad GEMMs, attention backward, LayerNorm backward (the dependency chain).  Side lane: wgrad GEMMs,
        bias column sums, LayerNorm partial reductions."""
        D, hid, H, hd = st.dim, st.hidden, st.heads, st.head_dim
        scale = hd ** -0.5
        cws, wsm, wss = pl.colsum_ws, self.ws_main, self.ws_side
        lanes = self.lanes
        d_d = pl.d_d[:M * D].view(M, D)
        delta = pl.delta[:B * H * N]
        for i in layers:
            pre = f"{st.prefix}.{i}"
            b, x_in = sb.blocks[i], sb.x[i]
            hb = i % self.ring_block
            self._prefetch_block(st, sb, i - 1, with_acts=True)
            d_hid = pl.d_hid[hb][:M * hid].view(M, hid)
            dqkv = pl.dqkv[hb][:M * 3 * D].view(M, 3 * D)
            dres16 = pl.dres16[cur][:M * D].view(M, D)
            # x_out = xmid + fc2(gelu(fc1(ln2))) + b2
            self._side(lambda: ops.gemm(dres16, b.act, D, hid, M, a_mn_major=True, b_mn_major=True,
                                        out_f32=self._g(f"{pre}.mlp.fc2.weight"), accumulate=acc, workspace=wss),
                       reads=(("dres16", cur),))
            lanes.before_write(("d_hid", hb))
            ops.gemm(dres16, self._w(f"{pre}.mlp.fc2.weight"), M, hid, D, b_mn_major=True, dgelu_src=b.pre,
                     out_bf16=d_hid, workspace=wsm)

            # the six column reductions of this block (two bias sums, two LayerNorm affine gradients, two residual sums)
            # are collected and run as ONE side-lane kernel at the end of the block
            jobs = [ops.col_job(d_hid, self._g(f"{pre}.mlp.fc1.bias"), cols=hid)]
            job_reads = [("d_hid", hb), ("dqkv", hb)]

            def side_fc1(d_hid=d_hid, b=b, pre=pre):
                ops.gemm(d_hid, b.ln2, hid, D, M, a_mn_major=True, b_mn_major=True,
                         out_f32=self._g(f"{pre}.mlp.fc1.weight"), accumulate=acc, workspace=wss)
            self._side(side_fc1, reads=(("d_hid", hb),))
            d_ln = self._ln_in(pl, M, D)
            ops.gemm(d_hid, self._w(f"{pre}.mlp.fc1.weight"), M, D, hid, b_mn_major=True, out_bf16=d_ln, workspace=wsm)
            nxt = (cur + 1) % self.ring
            self._ln_bwd(pl, d_ln, b.xmid, f"{pre}.norm2", b.mean2, b.rstd2, cur, nxt, M, D, acc,
                         f"{pre}.attn.proj.bias", jobs=jobs, reads=job_reads)
            cur = nxt
            dres16 = pl.dres16[cur][:M * D].view(M, D)
            # xmid = x_in + proj(attn(qkv(ln1))) + bp
            self._side(lambda dres16=dres16, b=b, pre=pre: ops.gemm(
                dres16, b.o, D, D, M, a_mn_major=True, b_mn_major=True, out_f32=self._g(f"{pre}.attn.proj.weight"),
                accumulate=acc, workspace=wss), reads=(("dres16", cur),))
            ops.gemm(dres16, self._w(f"{pre}.attn.proj.weight"), M, D, D, b_mn_major=True, out_bf16=d_d, workspace=wsm)
            lanes.before_write(("dqkv", hb))
            ops.attention_bwd(b.qkv, b.o, d_d, b.lse, delta, dqkv, B, N, H, hd, scale)

            jobs.append(ops.col_job(dqkv, self._g(f"{pre}.attn.qkv.bias"), cols=3 * D))

            def side_qkv(dqkv=dqkv, b=b, pre=pre):
                ops.gemm(dqkv, b.ln1, 3 * D, D, M, a_mn_major=True, b_mn_major=True,
                         out_f32=self._g(f"{pre}.attn.qkv.weight"), accumulate=acc, workspace=wss)
            self._side(side_qkv, reads=(("dqkv", hb),))
            d_ln = self._ln_in(pl, M, D)
            ops.gemm(dqkv, self._w(f"{pre}.attn.qkv.weight"), M, D, 3 * D, b_mn_major=True, out_bf16=d_ln, workspace=wsm)
            nxt = (cur + 1) % self.ring
            below = f"{st.prefix}.{i - 1}.mlp.fc2.bias" if i > 0 else None
            self._ln_bwd(pl, d_ln, x_in, f"{pre}.norm1", b.mean1, b.rstd1, cur, nxt, M, D, acc, below, jobs=jobs,
                         reads=job_reads)
            cur = nxt
            self._side(lambda jobs=jobs: ops.block_colreduce(jobs, M, pl.bcr_ws, accumulate=acc), reads=tuple(job_reads),
                       lane=1)
        return cur

    # ------------------------------------------------------------------------------------------------ contrastive predictor
    def _pred_bufs(self, pl: MAEPlan):
        if getattr(pl, "pb", None) is None:
            D, M, dev = self.enc.dim, pl.Me, self.device
            b = _BlockBufs()
            b.h = torch.empty((M, D), dtype=_F32, device=dev)
            b.act = torch.empty((M, D), dtype=_BF16, device=dev)
            b.mean, b.rstd = torch.empty(D, dtype=_F32, device=dev), torch.empty(D, dtype=_F32, device=dev)
            b.p = torch.empty((M, D), dtype=_F32, device=dev)
            b.dp32 = torch.empty((M, D), dtype=_F32, device=dev)
            b.dp16 = torch.empty((M, D), dtype=_BF16, device=dev)
            b.dact = torch.empty((M, D), dtype=_BF16, device=dev)
            b.dh = torch.empty((M, D), dtype=_BF16, device=dev)
            b.dlat = torch.empty((M, D), dtype=_F32, device=dev)
            b.cs_ws = torch.zeros(ops.colsum_workspace_bytes(M, D), dtype=torch.uint8, device=dev)
            pl.pb = b
        return pl.pb

    def predictor_forward(self, pl: MAEPlan) -> None:
        """p = predictor(latent) of model/vit_autoenc.py:263-268,282-283 on the rows of ``pl.latent`` (bf16 GEMM operand) ->
        pl.pb.p fp32 [B*Ne, D].  BatchNorm1d runs in training mode (batch statistics of the token rows, running statistics
        updated in place)."""
        b = self._pred_bufs(pl)
        D, M = self.enc.dim, pl.Me
        rm, rv, eps, mom = self.bn_state
        self._need("predictor.0.weight")
        ops.gemm(pl.latent, self._w("predictor.0.weight"), M, D, D, out_f32=b.h, workspace=self.ws_main)
        ops.bn_relu_fwd(b.h, self._p("predictor.1.weight"), self._p("predictor.1.bias"), eps, b.act, b.mean, b.rstd, rm, rv, mom)
        ops.gemm(b.act, self._w("predictor.3.weight"), M, D, D, bias=self._p("predictor.3.bias"), out_f32=b.p, workspace=self.ws_main)

    def predictor_backward(self, pl: MAEPlan, dp: torch.Tensor, accumulate: bool) -> torch.Tensor:
        """Reverse of predictor_forward for the upstream gradient ``dp`` fp32 [B*Ne, D]: parameter gradients into the flat
        buffer, returns the gradient w.r.t. the latent (fp32, a plan buffer)."""
        b = self._pred_bufs(pl)
        D, M = self.enc.dim, pl.Me
        b.dp32.copy_(dp.reshape(M, D))

        def body():
            ops.cast_f32_to_bf16(b.dp32.view(-1), b.dp16.view(-1))
            ops.gemm(b.dp16, b.act, D, D, M, a_mn_major=True, b_mn_major=True, out_f32=self._g("predictor.3.weight"),
                     accumulate=accumulate, workspace=self.ws_main)
            ops.colsum(b.dp16, M, D, self._g("predictor.3.bias"), b.cs_ws, accumulate=accumulate)
            ops.gemm(b.dp16, self._w("predictor.3.weight"), M, D, D, b_mn_major=True, out_bf16=b.dact, workspace=self.ws_main)
            ops.bn_relu_bwd(b.dact, b.h, self._p("predictor.1.weight"), self._p("predictor.1.bias"), b.mean, b.rstd, b.dh,
                            self._g("predictor.1.weight"), self._g("predictor.1.bias"), accumulate)
            ops.gemm(b.dh, pl.latent, D, D, M, a_mn_major=True, b_mn_major=True, out_f32=self._g("predictor.0.weight"),
                     accumulate=accumulate, workspace=self.ws_main)
            ops.gemm(b.dh, self._w("predictor.0.weight"), M, D, D, b_mn_major=True, out_f32=b.dlat, workspace=self.ws_main)
        self._run(pl, ("pred_bwd", bool(accumulate)), body)
        return b.dlat

    def zero_predictor_grads(self) -> None:
        a = self.flat.offsets[PREDICTOR_PARAMS[0]][0]
        o, k, _ = self.flat.offsets[PREDICTOR_PARAMS[-1]]
        self.flat.g32[a:o + k].zero_()

    # ------------------------------------------------------------------------------------------------ data parallel
    def _grad_reducer(self) -> "dp.GradReducer":
        """Reducer of the staged backward: bf16 exchange through a staging buffer on a communication stream of its own
        (dp.exchange_dtype), else plain fp32 all-reduces on the process group's stream."""
        if dp.exchange_dtype() != "bf16":
            return dp.GradReducer()
        if self.g16 is None:
            self.g16 = torch.empty(self.flat.total, dtype=_BF16, device=self.device)
            self.comm_stream = torch.cuda.Stream(device=self.device, priority=-1)
        return dp.GradReducer(self.flat.g32, self.g16, self.comm_stream)

    def broadcast_parameters(self) -> None:
        """One broadcast of the flat parameter buffer (+ the frozen position tables) from rank 0 (dp.py)."""
        dp.broadcast_flat(self.flat.p32)
        dp.broadcast_flat(self.pos)
        dp.broadcast_flat(self.dpos)

    def allreduce_gradients(self) -> None:
        """Mean of the whole flat gradient buffer over ranks in a few large slices (not overlapped; the training step uses
        backward(sync_grads=True), which overlaps the exchange with the backward stages)."""
        if dp.world_size() > 1:
            if self.grad_buckets is None:
                offs = [(self.flat.offsets[n][0], self.flat.offsets[n][1]) for n in self.flat.order]
                self.grad_buckets = dp.bucket_slices(offs, self.flat.total, self.bucket_elems)
            dp.allreduce_mean_bucketed_(self.flat.g32, self.grad_buckets)
        self.grads_local = False

    def sync_master(self) -> None:
        """After sharded optimizer steps the fp32 master of the large matrices is current only on its owner: complete
        this rank's copy (peer reads; called by the model's state_dict())."""
        if self.flat.sharded is not None:
            self.flat.sharded.sync_master()

    # ------------------------------------------------------------------------------------------------ optimizer
    def fused_optimizer(self) -> "FusedAdamW":
        if self.optim is None:
            self.optim = FusedAdamW(self)
        return self.optim


class FusedAdamW:
    """GradScaler.unscale_ + gradient norm + AdamW + GradScaler.update over the flat buffers in three launches
    (include/vitae_b200.h: vitae_optim_prepare / vitae_adamw_flat), driven by the hyper-parameters of the caller's own
    ``torch.optim.AdamW`` (k_fold_cross_valid_combined_brats.py:168-169).  The moments live in flat buffers; the
    optimizer's ``state[p]['exp_avg' / 'exp_avg_sq']`` are views of them, so ``optimizer.state_dict()`` checkpoints and a
    later plain ``optimizer.step()`` both keep working."""

    def __init__(self, eng: MAEEngine):
        flat, dev = eng.flat, eng.device
        self.eng = eng
        self.m = dp.alloc_flat(flat.total, _F32, dev, owner=flat)
        self.v = dp.alloc_flat(flat.total, _F32, dev, owner=flat)
        self._sharded_tried = False
        self.ctl = torch.zeros(8, dtype=_F32, device=dev)
        self.ws = torch.empty(ops.optim_workspace_bytes(), dtype=torch.uint8, device=dev)
        self.group_map = torch.full((flat.total // _ALIGN,), 255, dtype=torch.uint8, device=dev)
        self.bound: Optional[int] = None       # id of the bound optimizer
        self.bound_sig = None
        self.host_steps = 0
        # parameters of the optimizer that live outside the engine's flat buffers (contrastive predictor)
        self.ex_total, self.ex_items, self.ex_ptrs, self.ex_steps = 0, [], set(), []
        self.ex_p32 = self.ex_g32 = self.ex_m = self.ex_v = self.ex_group_map = None

    @staticmethod
    def supports(optimizer) -> bool:
        if type(optimizer) is not torch.optim.AdamW:
            return False
        return all(not g.get("amsgrad", False) and not g.get("maximize", False) and not g.get("capturable", False)
                   and not isinstance(g["lr"], torch.Tensor) for g in optimizer.param_groups)

    def bind(self, optimizer) -> bool:
        """Maps the optimizer's parameter groups onto 64-element chunks of the flat buffer and adopts / installs its
        state.  Parameters of the optimizer that are not flat views of this engine (the contrastive predictor) are moved
        into a second, small flat buffer of their own ("extras": their ``.data`` is re-pointed, values kept) and updated by
        the same kernel.  Returns False when the optimizer cannot be driven by the fused step."""
        flat = self.eng.flat
        by_ptr = {flat.v32[n].data_ptr(): n for n in flat.order}
        sig = tuple(tuple(p.data_ptr() for p in g["params"]) for g in optimizer.param_groups)
        if self.bound == id(optimizer) and self.bound_sig == sig and self._state_aliased(optimizer):
            return True
        if len(optimizer.param_groups) > 8:
            return False
        extras = [(gi, p) for gi, g in enumerate(optimizer.param_groups) for p in g["params"]
                  if p.data_ptr() not in by_ptr and p.data_ptr() not in self.ex_ptrs]
        if any((not p.is_cuda) or p.dtype != _F32 or p.device != flat.p32.device for _, p in extras):
            return False
        if extras or self.ex_total:
            self._build_extras(optimizer, by_ptr)
        gm = torch.full((flat.total // _ALIGN,), 255, dtype=torch.uint8)
        steps = []
        for gi, group in enumerate(optimizer.param_groups):
            for p in group["params"]:
                n = by_ptr.get(p.data_ptr())
                if n is None:
                    continue            # an extra: handled by _build_extras
                o, k, shp = flat.offsets[n]
                gm[o // _ALIGN:(o + k + _ALIGN - 1) // _ALIGN] = gi
                st = optimizer.state.get(p)
                mv, vv = self.m[o:o + k].view(shp), self.v[o:o + k].view(shp)
                if st and "exp_avg" in st:
                    if st["exp_avg"].data_ptr() != mv.data_ptr():
                        mv.copy_(st["exp_avg"]); vv.copy_(st["exp_avg_sq"])
                    steps.append(float(st["step"]))
                else:                   # a fresh optimizer starts from zero moments, whatever an earlier one left here
                    mv.zero_(); vv.zero_()
                optimizer.state[p] = {"step": torch.tensor(0.0), "exp_avg": mv, "exp_avg_sq": vv}
        steps += self.ex_steps
        self.group_map.copy_(gm)
        self.host_steps = int(max(steps)) if steps else 0
        self.ctl[5] = float(self.host_steps)
        for group in optimizer.param_groups:
            for p in group["params"]:
                optimizer.state[p]["step"].fill_(float(self.host_steps))
        if self.bound != id(optimizer):
            # optimizer.state_dict() (misc.save_model, reference misc.py:302) must see the device-side step count
            optimizer.register_state_dict_pre_hook(lambda opt, fo=weakref.ref(self): fo() and fo()._pre_state_dict(opt))
        self.bound = id(optimizer)
        self.bound_sig = tuple(tuple(p.data_ptr() for p in g["params"]) for g in optimizer.param_groups)
        return True

    def sharded(self) -> Optional["dp.ShardedStep"]:
        """The sharded data-parallel step (dp.ShardedStep), set up on first use -- a COLLECTIVE call: every rank must reach
        it (utils.misc.NativeScalerWithGradNormCount does, before the first backward of the fused path).  None when the job
        is not a single-node NCCL group of 2..8 ranks, the flat buffers are not symmetric allocations (engine built before
        init_process_group), or some optimizer parameters live outside the flat buffers."""
        eng, flat = self.eng, self.eng.flat
        if self._sharded_tried:
            return flat.sharded
        self._sharded_tried = True
        if not dp.sharded_enabled():
            return None
        ok = (getattr(eng, "supports_sharded", False) and not self.ex_total
              and all(dp.is_symmetric(t) for t in (flat.g32, flat.p32, flat.p16, self.m, self.v)))
        agree = torch.tensor([1.0 if ok else 0.0], device=eng.device)
        torch.distributed.all_reduce(agree, op=torch.distributed.ReduceOp.MIN)
        if agree.item() == 1.0:
            flat.sharded = dp.ShardedStep(flat, self.m, self.v, self.group_map)
        return flat.sharded

    def _state_aliased(self, optimizer) -> bool:
        """False once the optimizer's moments stopped being views of the flat buffers (``optimizer.load_state_dict``
        installs fresh tensors): bind() then adopts the loaded state."""
        flat = self.eng.flat
        n = flat.order[0]
        st = optimizer.state.get(flat.params[n])
        o = flat.offsets[n][0]
        return bool(st) and "exp_avg" in st and st["exp_avg"].data_ptr() == self.m[o:].data_ptr()

    def _pre_state_dict(self, optimizer) -> None:
        if self.bound == id(optimizer) and self._state_aliased(optimizer):
            self.sync_state(optimizer)

    def _build_extras(self, optimizer, by_ptr) -> None:
        """(Re)builds the extras flat buffers from every optimizer parameter that is not an engine view."""
        dev = self.eng.device
        items, off = [], 0
        for gi, group in enumerate(optimizer.param_groups):
            for p in group["params"]:
                if p.data_ptr() in by_ptr:
                    continue
                items.append((gi, p, off))
                off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        total = off
        p32, g32 = torch.zeros(total, dtype=_F32, device=dev), torch.zeros(total, dtype=_F32, device=dev)
        m, v = torch.zeros(total, dtype=_F32, device=dev), torch.zeros(total, dtype=_F32, device=dev)
        gm = torch.full((max(1, total // _ALIGN),), 255, dtype=torch.uint8)
        self.ex_steps, self.ex_items, self.ex_ptrs = [], [], set()
        with torch.no_grad():
            for gi, p, o in items:
                k, shp = p.numel(), p.shape
                pv = p32[o:o + k].view(shp)
                pv.copy_(p.data)
                st = optimizer.state.get(p)
                mv, vv = m[o:o + k].view(shp), v[o:o + k].view(shp)
                if st and "exp_avg" in st:
                    mv.copy_(st["exp_avg"]); vv.copy_(st["exp_avg_sq"])
                    self.ex_steps.append(float(st["step"]))
                p.data = pv
                optimizer.state[p] = {"step": torch.tensor(0.0), "exp_avg": mv, "exp_avg_sq": vv}
                gm[o // _ALIGN:(o + k + _ALIGN - 1) // _ALIGN] = gi
                self.ex_items.append((p, g32[o:o + k].view(shp)))
                self.ex_ptrs.add(pv.data_ptr())
        self.ex_p32, self.ex_g32, self.ex_m, self.ex_v, self.ex_total = p32, g32, m, v, total
        self.ex_group_map = gm.to(dev)

    def step(self, optimizer, scaler=None, overlap: bool = False) -> torch.Tensor:
        """One optimizer step on the gradients currently in the flat gradient buffer; returns the (unscaled) global
        gradient norm as a 0-dim device tensor.  ``scaler``: a torch.amp.GradScaler-like object whose scale lives in
        ``ctl[0]`` (see utils/misc.py::NativeScalerWithGradNormCount) or None.

        ``overlap``: the AdamW pass (HBM-bound: 30 bytes per parameter) is issued per parameter group, in forward order,
        on the engine's optimizer stream; the next ``forward`` waits for each group right before its first use, so the
        update of the later layers hides behind the forward of the earlier ones.  Until that forward (or
        ``engine.wait_params()``) the caller's stream is NOT ordered after the update."""
        eng, flat = self.eng, self.eng.flat
        use_scaler = scaler is not None
        gf, bf, gi = (scaler.get_growth_factor(), scaler.get_backoff_factor(), scaler.get_growth_interval()) \
            if use_scaler else (2.0, 0.5, 2000)
        eng.wait_params()
        if getattr(eng, "grads_local", False):
            sh = flat.sharded
            if sh is None or self.ex_total:
                eng.allreduce_gradients()        # the exchange was deferred to a sharded step that cannot run
            else:
                rows = [(g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"]) for g in optimizer.param_groups]
                lap = (eng.opt_stream, eng.group_ranges, eng.param_ready) if sh.overlap_gather and hasattr(eng, "param_ready") \
                    else None
                if sh.step(self.ctl, rows, gf, bf, gi, use_scaler, overlap=lap):
                    eng.params_in_flight = True
                eng.grads_local = False
                eng.norm_partials = 0
                norm = self.ctl[4].clone()
                flat.stamp_shadow()
                flat.overwrite_grads = True
                self.host_steps += 1
                return norm
        if self.ex_total:      # gradients of the extras come from torch autograd: gather them into their flat buffer
            have = [(dst, p.grad) for p, dst in self.ex_items if p.grad is not None]
            if len(have) != len(self.ex_items):
                self.ex_g32.zero_()
            if have:
                torch._foreach_copy_([d for d, _ in have], [g for _, g in have])
        if eng.norm_partials and not self.ex_total:
            # the backward left the squared-norm partials of every gradient slice behind (computed under its later stages)
            ops.optim_finalize(eng.norm_ws, eng.norm_partials, self.ctl, gf, bf, gi, use_scaler)
        else:
            ops.optim_prepare(flat.g32, flat.total, self.ctl, self.ws, gf, bf, gi, use_scaler,
                              grad2=self.ex_g32 if self.ex_total else None, n2=self.ex_total)
        eng.norm_partials = 0
        rows = [(g["lr"], g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"]) for g in optimizer.param_groups]
        norm = self.ctl[4].clone()
        if self.ex_total:
            ops.adamw_flat(self.ex_p32, self.ex_g32, self.ex_m, self.ex_v, None, self.ex_total, self.ex_group_map, rows,
                           self.ctl)
        if overlap:
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(main)
            eng.opt_stream.wait_event(ev)
            with torch.cuda.stream(eng.opt_stream):
                for g, (a, b) in enumerate(eng.group_ranges):          # forward order
                    ops.adamw_flat(flat.p32, flat.g32, self.m, self.v, flat.p16, b - a, self.group_map, rows, self.ctl,
                                   start=a, max_blocks=148 * 2)
                    eng.param_ready[g].record(eng.opt_stream)
            eng.params_in_flight = True
        else:
            ops.adamw_flat(flat.p32, flat.g32, self.m, self.v, flat.p16, flat.total, self.group_map, rows, self.ctl)
        flat.stamp_shadow()
        flat.overwrite_grads = True      # these gradients are consumed: the next backward starts from zero
        self.host_steps += 1
        return norm

    def sync_state(self, optimizer) -> None:
        """Writes the device-side step count (skipped steps excluded) into the optimizer's per-parameter state (one
        device read): call before ``optimizer.state_dict()`` / switching to ``optimizer.step()``."""
        self.eng.wait_params()
        if self.eng.flat.sharded is not None:      # moments of the parts other ranks own
            self.eng.flat.sharded.sync_moments()
        steps = float(self.ctl[5].item())
        self.host_steps = int(steps)
        for group in optimizer.param_groups:
            for p in group["params"]:
                st = optimizer.state.get(p)
                if st is not None and "step" in st:
                    st["step"].fill_(steps)
