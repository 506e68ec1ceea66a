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

Mixed precision: bf16 tensor-core operands, fp32 accumulation, fp32 residual stream / LayerNorm statistics /
loss / gradients of parameters (BASELINE.json north_star tolerance for this mode: 1e-2 relative).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import ops

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


def backward_param_order(depth: int, decoder_depth: int) -> List[str]:
    names = ["decoder_pred.weight", "decoder_pred.bias", "decoder_norm.weight", "decoder_norm.bias"]
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
        self.p32 = torch.zeros(off, dtype=_F32, device=device)
        self.p16 = torch.zeros(off, dtype=_BF16, device=device)
        self.g32 = torch.zeros(off, dtype=_F32, device=device)
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
        self.shadow_fresh = False

    def still_aliased(self) -> bool:
        n = self.order[0]
        return self.params[n].data_ptr() == self.v32[n].data_ptr() and self.params[n].device == self.p32.device

    def refresh_shadow(self) -> None:
        """fp32 master -> bf16 shadow (one launch).  Skipped when the fused optimizer already wrote the shadow."""
        if not self.shadow_fresh:
            ops.cast_params_bf16(self.cast_table, 1, self.p16, self.total)

    def grads_alias(self) -> Optional[bool]:
        """True: every .grad is our view (accumulate in place); False: every .grad is None; None: mixed / foreign."""
        state = None
        for n in self.order:
            g = self.params[n].grad
            s = False if g is None else (True if g.data_ptr() == self.vg[n].data_ptr() else None)
            if s is None:
                return None
            if state is None:
                state = s
            elif state != s:
                return None
        return state


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
        self.mean_n, self.rstd_n = a.new((self.Me,), _F32), a.new((self.Me,), _F32)
        self.hN = a.new((self.Md, Dd), _BF16)
        self.mean_dn, self.rstd_dn = a.new((self.Md,), _F32), a.new((self.Md,), _F32)
        self.pred = a.new((B, self.Nd, P), _BF16)
        self.pred32: Optional[torch.Tensor] = None   # fp32 copy of pred, allocated when the caller wants fp32 back
        self.step_id = 0
        self.patch_sums = a.new((B * L,), _F32)
        self.loss_out = a.new((2,), _F32)
        # backward scratch
        Mmax = max(self.Me, self.Md)
        Dmax = max(D, Dd)
        Hmax = max(eng.enc.hidden, eng.dec.hidden)
        self.dloss = a.new((1,), _F32)
        self.dpred = a.new((B, self.Nd, P), _BF16)
        self.dres = [a.new((Mmax * Dmax,), _F32), a.new((Mmax * Dmax,), _F32)]
        self.dres16 = a.new((Mmax * Dmax,), _BF16)
        self.d_d = a.new((Mmax * Dmax,), _BF16)
        self.d_hid = a.new((Mmax * Hmax,), _BF16)
        self.dqkv = a.new((Mmax * 3 * Dmax,), _BF16)
        self.delta = a.new((B * max(eng.enc.heads * self.Ne, eng.dec.heads * self.Nd),), _F32)
        nb = ops.layernorm_bwd_blocks(Mmax)
        self.ln_partials = a.new((2 * nb * Dmax,), _F32)
        cs_rows = max(Mmax, nb)
        self.colsum_ws = a.new((ops.colsum_blocks(cs_rows) * max(P, 3 * Dmax, Hmax),), _F32)
        self.g_embed = a.new((self.Me, Dd), _BF16)
        self.g_pe = a.new((B * keep, D), _BF16)
        self.nbytes = a.nbytes
        self.vol: Optional[torch.Tensor] = None   # the caller's volume of the current step (read in place)

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
        for st in (self.enc, self.dec):
            if st.head_dim not in (16, 32, 64):
                raise ops._lib.VitaeError(f"unsupported head_dim {st.head_dim} (kernels exist for 16/32/64)")
            if st.dim % 8 or st.hidden % 8 or st.dim > 1024:
                raise ops._lib.VitaeError(f"unsupported width {st.dim}/{st.hidden}")
        if self.P % 8 or p % 4:
            raise ops._lib.VitaeError("patch_size must be a multiple of 4")
        self.flat = FlatParams(named_params, backward_param_order(cfg["depth"], cfg["decoder_depth"]), dev)
        self.pos = pos_embed.detach().reshape(self.L + 1, D).contiguous()
        self.dpos = decoder_pos_embed.detach().reshape(self.L + 1, Dd).contiguous()
        self.plans: Dict[Tuple[int, int], MAEPlan] = {}
        self.kernel_launches = 0

    # ------------------------------------------------------------------------------------------------ helpers
    def plan(self, B: int, keep: int) -> MAEPlan:
        key = (B, keep)
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

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, vol: torch.Tensor, noise: torch.Tensor, keep: int, want_loss: bool = True,
                pred_f32: bool = False) -> MAEPlan:
        """vol fp32 [B,C,V,V,V] (contiguous, CUDA); noise fp32 [B,L].  Fills plan.pred / mask / loss_out."""
        B = vol.shape[0]
        pl = self.plan(B, keep)
        pl.vol = vol
        self.flat.refresh_shadow()
        self.encode(pl, vol, noise)
        self.decode(pl, pred_f32)
        if want_loss:
            ops.masked_mse_fwd(pl.pred, vol, pl.mask, pl.patch_sums, pl.loss_out, self.p)
        return pl

    def encode(self, pl: MAEPlan, vol: torch.Tensor, noise: torch.Tensor) -> None:
        """model/vit_autoenc.py:157-177 (forward_encoder): patch embed of the kept patches only (the dropped ones are
        discarded by the gather at :147, so embedding them is dead work) + pos, cls row, blocks, final norm."""
        B, keep, D = pl.B, pl.keep, self.enc.dim
        ops.random_masking(noise, pl.ids_shuffle, pl.ids_restore, pl.mask, keep)
        ops.build_row_maps(pl.ids_shuffle, keep, pl.maps)
        ops.im2col_patches(vol, pl.ids_shuffle, pl.cols, self.p, keep)
        x0 = pl.enc.x[0]
        ops.gemm(pl.cols, self._w("patch_embed.proj.weight"), B * keep, D, self.Kpe,
                 bias=self._p("patch_embed.proj.bias"), addend=self.pos, add_rows=pl.maps["pe_pos_rows"], ldadd=D,
                 out_f32=x0, out_rows=pl.maps["enc_tok_rows"])
        ops.fill_rows(x0, pl.maps["enc_cls_rows"], B, D, self._p("cls_token"), None, self.pos, None)
        self._stack_fwd(self.enc, pl.enc, pl.Me, B, pl.Ne)
        ops.layernorm_fwd(pl.enc.x[-1], self._p("norm.weight"), self._p("norm.bias"), pl.latent, pl.mean_n, pl.rstd_n,
                          self.eps)

    def decode(self, pl: MAEPlan, pred_f32: bool = False) -> None:
        """model/vit_autoenc.py:179-203 (forward_decoder); pl.pred keeps the cls row (row 0 of each sample)."""
        if pred_f32 and pl.pred32 is None:
            pl.pred32 = torch.empty((pl.B, pl.Nd, self.P), dtype=_F32, device=self.device)
        B, D, Dd = pl.B, self.enc.dim, self.dec.dim
        xd0 = pl.dec.x[0]
        ops.gemm(pl.latent, self._w("decoder_embed.weight"), pl.Me, Dd, D, bias=self._p("decoder_embed.bias"),
                 addend=self.dpos, add_rows=pl.maps["dec_pos_rows_of_enc"], ldadd=Dd, out_f32=xd0,
                 out_rows=pl.maps["dec_rows_of_enc"])
        if pl.nmask > 0:
            ops.fill_rows(xd0, pl.maps["masked_dec_rows"], B * pl.nmask, Dd, self._p("mask_token"), None, self.dpos,
                          pl.maps["masked_pos_rows"])
        self._stack_fwd(self.dec, pl.dec, pl.Md, B, pl.Nd)
        ops.layernorm_fwd(pl.dec.x[-1], self._p("decoder_norm.weight"), self._p("decoder_norm.bias"), pl.hN,
                          pl.mean_dn, pl.rstd_dn, self.eps)
        ops.gemm(pl.hN, self._w("decoder_pred.weight"), pl.Md, self.P, Dd, bias=self._p("decoder_pred.bias"),
                 out_bf16=pl.pred.view(pl.Md, self.P), out_f32=pl.pred32.view(pl.Md, self.P) if pred_f32 else None)

    def _stack_fwd(self, st: StackSpec, sb, M: int, B: int, N: int) -> None:
        """model/vit.py:139-144 (Block), :112-124 (Attention), :90-96 (Mlp3D)."""
        D, hid, H, hd = st.dim, st.hidden, st.heads, st.head_dim
        scale = hd ** -0.5
        for i in range(st.depth):
            pre = f"{st.prefix}.{i}"
            b, x_in, x_out = sb.blocks[i], sb.x[i], sb.x[i + 1]
            ops.layernorm_fwd(x_in, self._p(f"{pre}.norm1.weight"), self._p(f"{pre}.norm1.bias"), b.ln1, b.mean1,
                              b.rstd1, self.eps)
            ops.gemm(b.ln1, self._w(f"{pre}.attn.qkv.weight"), M, 3 * D, D, bias=self._p(f"{pre}.attn.qkv.bias"),
                     out_bf16=b.qkv)
            ops.attention_fwd(b.qkv, b.o, b.lse, B, N, H, hd, scale)
            ops.gemm(b.o, self._w(f"{pre}.attn.proj.weight"), M, D, D, bias=self._p(f"{pre}.attn.proj.bias"),
                     addend=x_in, out_f32=b.xmid)
            ops.layernorm_fwd(b.xmid, self._p(f"{pre}.norm2.weight"), self._p(f"{pre}.norm2.bias"), b.ln2, b.mean2,
                              b.rstd2, self.eps)
            ops.gemm(b.ln2, self._w(f"{pre}.mlp.fc1.weight"), M, hid, D, bias=self._p(f"{pre}.mlp.fc1.bias"),
                     out_bf16=b.pre, out_gelu_bf16=b.act)
            ops.gemm(b.act, self._w(f"{pre}.mlp.fc2.weight"), M, D, hid, bias=self._p(f"{pre}.mlp.fc2.bias"),
                     addend=b.xmid, out_f32=x_out)

    # ------------------------------------------------------------------------------------------------ backward
    def backward(self, pl: MAEPlan, dloss: Optional[torch.Tensor], dpred_extra: Optional[torch.Tensor] = None,
                 accumulate: bool = False) -> None:
        """Gradient of (dloss * recon_loss [+ <dpred_extra, pred>]) w.r.t. every trainable parameter, written to
        (accumulate=False) or added into (True) the flat gradient buffer.  Hand-derived reverse of forward()."""
        B, D, Dd, P = pl.B, self.enc.dim, self.dec.dim, self.P
        acc = accumulate
        ws = pl.colsum_ws
        if dloss is None:
            pl.dloss.zero_()
        else:
            pl.dloss.copy_(dloss.reshape(1))
        # ---- loss: d recon / d pred (model/vit_autoenc.py:226-227), zeros for kept patches and the cls row
        ops.masked_mse_bwd(pl.pred, pl.vol, pl.mask, pl.loss_out[1:], pl.dloss, pl.dpred, self.p)
        if dpred_extra is not None:   # gradient of auxiliary torch-side terms that consume ``pred`` (edge-map loss)
            pl.dpred[:, 1:, :].add_(dpred_extra.to(_BF16))
        dpred = pl.dpred.view(pl.Md, P)
        # ---- decoder_pred (vit_autoenc.py:198)
        ops.gemm(dpred, pl.hN, P, Dd, pl.Md, a_mn_major=True, b_mn_major=True, out_f32=self._g("decoder_pred.weight"),
                 accumulate=acc)
        ops.colsum(dpred, pl.Md, P, self._g("decoder_pred.bias"), ws, accumulate=acc)
        d_d = pl.d_d[:pl.Md * Dd].view(pl.Md, Dd)
        ops.gemm(dpred, self._w("decoder_pred.weight"), pl.Md, Dd, P, b_mn_major=True, out_bf16=d_d)
        cur = self._ln_bwd(pl, d_d, pl.dec.x[-1], "decoder_norm", pl.mean_dn, pl.rstd_dn, None, 0, pl.Md, Dd, acc)
        cur = self._stack_bwd(self.dec, pl.dec, pl, pl.Md, B, pl.Nd, cur, acc)
        dxd = pl.dres[cur][:pl.Md * Dd].view(pl.Md, Dd)
        # ---- mask tokens, decoder_embed (vit_autoenc.py:181-190)
        if pl.nmask > 0:
            ops.sum_rows(dxd, pl.maps["masked_dec_rows"], B * pl.nmask, Dd, self._g("mask_token").view(-1), acc)
        elif not acc:
            self._g("mask_token").zero_()
        ops.gather_rows(dxd, pl.maps["dec_rows_of_enc"], pl.Me, Dd, pl.g_embed, None)
        ops.gemm(pl.g_embed, pl.latent, Dd, D, pl.Me, a_mn_major=True, b_mn_major=True,
                 out_f32=self._g("decoder_embed.weight"), accumulate=acc)
        ops.colsum(pl.g_embed, pl.Me, Dd, self._g("decoder_embed.bias"), ws, accumulate=acc)
        d_e = pl.d_d[:pl.Me * D].view(pl.Me, D)
        ops.gemm(pl.g_embed, self._w("decoder_embed.weight"), pl.Me, D, Dd, b_mn_major=True, out_bf16=d_e)
        # ---- encoder norm + blocks (vit_autoenc.py:172-175)
        cur = self._ln_bwd(pl, d_e, pl.enc.x[-1], "norm", pl.mean_n, pl.rstd_n, None, 0, pl.Me, D, acc)
        cur = self._stack_bwd(self.enc, pl.enc, pl, pl.Me, B, pl.Ne, cur, acc)
        dx0 = pl.dres[cur][:pl.Me * D].view(pl.Me, D)
        # ---- cls token, patch embed (vit_autoenc.py:160-170); no input gradient for the volume
        ops.sum_rows(dx0, pl.maps["enc_cls_rows"], B, D, self._g("cls_token").view(-1), acc)
        ops.gather_rows(dx0, pl.maps["enc_tok_rows"], B * pl.keep, D, pl.g_pe, None)
        ops.gemm(pl.g_pe, pl.cols, D, self.Kpe, B * pl.keep, a_mn_major=True, b_mn_major=True,
                 out_f32=self._g("patch_embed.proj.weight").view(D, self.Kpe), accumulate=acc)
        ops.colsum(pl.g_pe, B * pl.keep, D, self._g("patch_embed.proj.bias"), ws, accumulate=acc)

    def _ln_bwd(self, pl: MAEPlan, dy: torch.Tensor, x: torch.Tensor, name: str, mean, rstd, dx_in_idx: Optional[int],
                out_idx: int, M: int, D: int, acc: bool) -> int:
        """LayerNorm backward: dres[out_idx] = (dres[dx_in_idx] if given) + LN'(dy); refreshes the bf16 copy dres16 and
        the affine gradients.  Returns out_idx."""
        nb = ops.layernorm_bwd_blocks(M)
        partials = pl.ln_partials[:2 * nb * D].view(2, nb, D)
        dx_in = None if dx_in_idx is None else pl.dres[dx_in_idx][:M * D].view(M, D)
        dx_out = pl.dres[out_idx][:M * D].view(M, D)
        dx16 = pl.dres16[:M * D].view(M, D)
        ops.layernorm_bwd(dy, x, self._p(f"{name}.weight"), mean, rstd, dx_in, dx_out, dx16, partials)
        ops.colsum(partials[0], nb, D, self._g(f"{name}.weight"), pl.colsum_ws, accumulate=acc)
        ops.colsum(partials[1], nb, D, self._g(f"{name}.bias"), pl.colsum_ws, accumulate=acc)
        return out_idx

    def _stack_bwd(self, st: StackSpec, sb, pl: MAEPlan, M: int, B: int, N: int, cur: int, acc: bool) -> int:
        """Reverse of _stack_fwd.  On entry dres[cur] / dres16 hold the gradient w.r.t. the stack output."""
        D, hid, H, hd = st.dim, st.hidden, st.heads, st.head_dim
        scale = hd ** -0.5
        ws = pl.colsum_ws
        dres16 = pl.dres16[:M * D].view(M, D)
        d_hid = pl.d_hid[:M * hid].view(M, hid)
        d_d = pl.d_d[:M * D].view(M, D)
        dqkv = pl.dqkv[:M * 3 * D].view(M, 3 * D)
        delta = pl.delta[:B * H * N]
        for i in reversed(range(st.depth)):
            pre = f"{st.prefix}.{i}"
            b, x_in = sb.blocks[i], sb.x[i]
            dres = pl.dres[cur][:M * D].view(M, D)
            # x_out = xmid + fc2(gelu(fc1(ln2))) + b2
            ops.gemm(dres16, b.act, D, hid, M, a_mn_major=True, b_mn_major=True, out_f32=self._g(f"{pre}.mlp.fc2.weight"),
                     accumulate=acc)
            ops.colsum(dres, M, D, self._g(f"{pre}.mlp.fc2.bias"), ws, accumulate=acc)
            ops.gemm(dres16, self._w(f"{pre}.mlp.fc2.weight"), M, hid, D, b_mn_major=True, dgelu_src=b.pre, out_bf16=d_hid)
            ops.gemm(d_hid, b.ln2, hid, D, M, a_mn_major=True, b_mn_major=True, out_f32=self._g(f"{pre}.mlp.fc1.weight"),
                     accumulate=acc)
            ops.colsum(d_hid, M, hid, self._g(f"{pre}.mlp.fc1.bias"), ws, accumulate=acc)
            ops.gemm(d_hid, self._w(f"{pre}.mlp.fc1.weight"), M, D, hid, b_mn_major=True, out_bf16=d_d)
            nxt = cur ^ 1
            self._ln_bwd(pl, d_d, b.xmid, f"{pre}.norm2", b.mean2, b.rstd2, cur, nxt, M, D, acc)
            cur = nxt
            dres = pl.dres[cur][:M * D].view(M, D)
            # xmid = x_in + proj(attn(qkv(ln1))) + bp
            ops.gemm(dres16, b.o, D, D, M, a_mn_major=True, b_mn_major=True, out_f32=self._g(f"{pre}.attn.proj.weight"),
                     accumulate=acc)
            ops.colsum(dres, M, D, self._g(f"{pre}.attn.proj.bias"), ws, accumulate=acc)
            ops.gemm(dres16, self._w(f"{pre}.attn.proj.weight"), M, D, D, b_mn_major=True, out_bf16=d_d)
            ops.attention_bwd(b.qkv, b.o, d_d, b.lse, delta, dqkv, B, N, H, hd, scale)
            ops.gemm(dqkv, b.ln1, 3 * D, D, M, a_mn_major=True, b_mn_major=True, out_f32=self._g(f"{pre}.attn.qkv.weight"),
                     accumulate=acc)
            ops.colsum(dqkv, M, 3 * D, self._g(f"{pre}.attn.qkv.bias"), ws, accumulate=acc)
            ops.gemm(dqkv, self._w(f"{pre}.attn.qkv.weight"), M, D, 3 * D, b_mn_major=True, out_bf16=d_d)
            nxt = cur ^ 1
            self._ln_bwd(pl, d_d, x_in, f"{pre}.norm1", b.mean1, b.rstd1, cur, nxt, M, D, acc)
            cur = nxt
        return cur

    # ------------------------------------------------------------------------------------------------ data parallel
    @staticmethod
    def _world() -> int:
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def broadcast_parameters(self) -> None:
        """Ranks seed differently (k_fold_cross_valid_combined_brats.py:87) and the scripts never wrap DDP
        (:154), so the module replicates rank 0's parameters itself: one broadcast of the flat buffer."""
        if self._world() > 1:
            import torch.distributed as dist
            dist.broadcast(self.flat.p32, src=0)
            dist.broadcast(self.pos, src=0)
            dist.broadcast(self.dpos, src=0)

    def allreduce_gradients(self) -> None:
        """Mean of the flat gradient buffer over ranks (one exchange step per optimizer step)."""
        if self._world() > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat.g32, op=dist.ReduceOp.AVG)
