"""Thin torch-tensor front end over the C ABI (include/vitae_b200.h).  torch is only the allocator / stream
provider here: every function enqueues one hand-written sm_100a kernel (or a fixed short sequence) on the current
CUDA stream.  No fallbacks: a non-CUDA tensor or a missing library raises."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from . import _lib
from ._lib import GemmEpilogue, check

_BF16 = torch.bfloat16
_F32 = torch.float32
_I32 = torch.int32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise _lib.VitaeError(f"{name}: expected a CUDA tensor (this package has no CPU path)")
    if t.dtype != dtype:
        raise _lib.VitaeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.VitaeError(f"{name}: expected a contiguous tensor")


def gemm_config(M: int, N: int, K: int, n_sm: int = 148):
    """Heuristic (tile_n, split_k): fill ~1 wave of 148 SMs; split K only for long reductions on a small output grid
    (patch-embed forward, decoder_pred dgrad: K = 16384), where the slab pass costs less than the idle SMs."""
    tiles_m = (M + 127) // 128
    num_kb = (K + 63) // 64
    tile_n = 128
    if tiles_m * ((N + 127) // 128) < n_sm and N % 64 == 0:
        tile_n = 64
    tiles = tiles_m * ((N + tile_n - 1) // tile_n)
    split = 1
    if num_kb >= 32 and tiles <= n_sm:
        split = min(max(1, (2 * n_sm) // tiles), max(1, num_kb // 8), 16)
    return tile_n, split


# Measured (tile_n, split_k) per GEMM signature.  The GEMMs of this path are short and latency / ingest bound
# (DESIGN.md), so the best tiling is found by timing the candidates once per shape on the device (first eager call,
# vitae_gemm_autotune, ~2 ms per shape) instead of a static rule.  VITAE_GEMM_AUTOTUNE=0 keeps the static rule (run-to-run reproducible bits:
# split-K changes the summation order).
_tuned = {}
AUTOTUNE = os.environ.get("VITAE_GEMM_AUTOTUNE", "1") != "0"
# VITAE_GEMM_AUTOTUNE_COLD=1 times the candidates on a flushed L2 instead of back to back (measured: same step time, 4.65 vs 4.62 ms)
AUTOTUNE_COLD = os.environ.get("VITAE_GEMM_AUTOTUNE_COLD", "0") != "0"
_flush = {}


def _flush_buffer(device):
    """256 MB scratch (2x the B200's L2) overwritten before every timed launch of the autotuner; freed by
    release_autotune_scratch()."""
    buf = _flush.get(device.index)
    if buf is None:
        buf = _flush[device.index] = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    return buf


def release_autotune_scratch() -> None:
    _flush.clear()


class GrowBuf:
    """A device byte buffer that grows on demand (never inside a CUDA-graph capture: the eager warm-up sizes it).
    Superseded allocations are kept alive: CUDA graphs captured against the smaller buffer have its address baked in
    and keep replaying into it (split-K slabs are scratch that lives from a GEMM to its finalize kernel, so the old and
    the new buffer never need to agree), and freeing it would hand that memory to the caching allocator."""

    def __init__(self, device):
        self.device = device
        self.buf: Optional[torch.Tensor] = None
        self.retired: list = []

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes:
            if torch.cuda.is_current_stream_capturing():
                raise _lib.VitaeError("workspace would have to grow during CUDA-graph capture")
            if self.buf is not None:
                self.retired.append(self.buf)
            self.buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=self.device)
        return self.buf


_gemm_ws = {}
_gemm_log = None   # when a list: every gemm() call appends (flops, closure) -- bench.py's per-kernel roofline replay


class record_gemms:
    """Context manager: records the GEMM launches made inside it so that they can be replayed in isolation."""

    def __init__(self):
        self.calls = []

    def __enter__(self):
        global _gemm_log
        _gemm_log = self.calls
        return self

    def __exit__(self, *exc):
        global _gemm_log
        _gemm_log = None

    @property
    def flops(self) -> float:
        return float(sum(f for f, _ in self.calls))

    def replay(self) -> None:
        for _, fn in self.calls:
            fn()


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    ws = _gemm_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _gemm_ws[key] = ws
    return ws


def gemm(a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, *, a_mn_major: bool = False,
         b_mn_major: bool = False, lda: Optional[int] = None, ldb: Optional[int] = None,
         bias: Optional[torch.Tensor] = None, addend: Optional[torch.Tensor] = None,
         add_rows: Optional[torch.Tensor] = None, ldadd: Optional[int] = None,
         dgelu_src: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None,
         ld_f32: Optional[int] = None, accumulate: bool = False, out_bf16: Optional[torch.Tensor] = None,
         out_gelu_bf16: Optional[torch.Tensor] = None, ld_bf16: Optional[int] = None,
         out_rows: Optional[torch.Tensor] = None, alpha: float = 1.0, alpha_ptr: Optional[torch.Tensor] = None,
         tile_n: Optional[int] = None, split_k: Optional[int] = None, workspace: Optional[GrowBuf] = None) -> None:
    """acc[m,n] = sum_k A(m,k) B(n,k) on tcgen05 + fused epilogue; see include/vitae_b200.h."""
    lib = _lib.load()
    _req(a, _BF16, "gemm A"); _req(b, _BF16, "gemm B")
    if lda is None:
        lda = M if a_mn_major else K
    if ldb is None:
        ldb = N if b_mn_major else K
    ep = GemmEpilogue()
    ep.alpha = alpha
    ep.alpha_ptr = _ptr(alpha_ptr)
    ep.bias = _ptr(bias)
    ep.addend = _ptr(addend)
    ep.add_rows = _ptr(add_rows)
    ep.ldadd = ldadd if ldadd is not None else N
    ep.dgelu_src = _ptr(dgelu_src)
    ep.ld_dgelu = N
    ep.out_f32 = _ptr(out_f32)
    ep.ld_f32 = ld_f32 if ld_f32 is not None else N
    ep.accumulate = 1 if accumulate else 0
    ep.out_bf16 = _ptr(out_bf16)
    ep.out_gelu_bf16 = _ptr(out_gelu_bf16)
    ep.ld_bf16 = ld_bf16 if ld_bf16 is not None else N
    ep.out_rows = _ptr(out_rows)
    def launch_with(tn: int, sk: int):
        ws_ptr = None
        ws_bytes = lib.vitae_gemm_workspace_bytes_for(ctypes.byref(ep), int(a_mn_major), int(b_mn_major), M, N, sk)
        if ws_bytes:   # split-K slabs, or an epilogue that runs in the finalize kernel (row maps, unusual output sets)
            ws = workspace.get(ws_bytes) if workspace is not None else _workspace(ws_bytes, a.device)
            ws_ptr = ws.data_ptr()
        check(lib.vitae_gemm_bf16(a.data_ptr(), lda, int(a_mn_major), b.data_ptr(), ldb, int(b_mn_major), M, N, K,
                                  ctypes.byref(ep), tn, sk, ws_ptr, ws_bytes, _stream()), "vitae_gemm_bf16")

    if tile_n is None or split_k is None:
        key = (M, N, K, bool(a_mn_major), bool(b_mn_major), bias is not None, addend is not None, add_rows is not None,
               dgelu_src is not None, out_f32 is not None, out_bf16 is not None, out_gelu_bf16 is not None,
               out_rows is not None)
        cfg = _tuned.get(key)
        if cfg is None:
            if AUTOTUNE and not accumulate and not torch.cuda.is_current_stream_capturing():
                ws_bytes = max(lib.vitae_gemm_workspace_bytes_for(ctypes.byref(ep), int(a_mn_major), int(b_mn_major), M, N, sk)
                               for sk in (1, 8))
                ws = (workspace.get(ws_bytes) if workspace is not None else _workspace(ws_bytes, a.device)) if ws_bytes else None
                tn, sk = ctypes.c_int(0), ctypes.c_int(1)
                fl = _flush_buffer(a.device) if AUTOTUNE_COLD else None
                check(lib.vitae_gemm_autotune(a.data_ptr(), lda, int(a_mn_major), b.data_ptr(), ldb, int(b_mn_major), M, N, K,
                                              ctypes.byref(ep), _ptr(ws), ws_bytes, _ptr(fl), fl.numel() if fl is not None else 0,
                                              _stream(), ctypes.byref(tn), ctypes.byref(sk)), "vitae_gemm_autotune")
                cfg = _tuned[key] = (tn.value, sk.value)
            else:
                cfg = gemm_config(M, N, K)
        tile_n = cfg[0] if tile_n is None else tile_n
        split_k = cfg[1] if split_k is None else split_k

    def launch():
        launch_with(tile_n, split_k)
    launch()
    if _gemm_log is not None:
        _gemm_log.append((2.0 * M * N * K, launch))


def layernorm_fwd(x, gamma, beta, y_bf16, mean, rstd, eps: float, y_f32=None) -> None:
    lib = _lib.load()
    _req(x, _F32, "layernorm x")
    rows, D = x.numel() // x.shape[-1], x.shape[-1]
    check(lib.vitae_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(y_bf16), _ptr(y_f32),
                                  _ptr(mean), _ptr(rstd), rows, D, eps, _stream()), "vitae_layernorm_fwd")


def layernorm_bwd_blocks(rows: int) -> int:
    return _lib.load().vitae_layernorm_bwd_blocks(rows)


def colsum_workspace_bytes(rows: int, cols: int) -> int:
    return _lib.load().vitae_colsum_workspace_bytes(rows, cols)


def reduce_partials(partials, nblk: int, D: int, out0=None, out1=None, out2=None, accumulate: bool = False) -> None:
    """Finishes layernorm_bwd's partials [3, nblk, D] -> (dgamma, dbeta, column sums of dx_out); None outputs skipped."""
    lib = _lib.load()
    check(lib.vitae_reduce_partials(partials.data_ptr(), nblk, D, _ptr(out0), _ptr(out1), _ptr(out2), int(accumulate),
                                    _stream()), "vitae_reduce_partials")


def _dy_pair(dy, dy2):
    """(bf16 pointer, fp32 pointer) for one or two upstream gradients of different dtype (summed by the kernels)."""
    ptr = {_BF16: None, _F32: None}
    for t in (dy, dy2):
        if t is not None:
            if ptr[t.dtype] is not None:
                raise _lib.VitaeError("layernorm backward: two upstream gradients need different dtypes (bf16 + fp32)")
            ptr[t.dtype] = t.data_ptr()
    return ptr[_BF16], ptr[_F32]


def layernorm_bwd(dy, x, gamma, mean, rstd, dx_in, dx_out, dx_out_bf16, dy2=None) -> None:
    lib = _lib.load()
    rows, D = x.numel() // x.shape[-1], x.shape[-1]
    dy16, dy32 = _dy_pair(dy, dy2)
    check(lib.vitae_layernorm_bwd(dy16, dy32, x.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                  _ptr(dx_in), dx_out.data_ptr(), _ptr(dx_out_bf16), rows, D, _stream()),
          "vitae_layernorm_bwd")


def layernorm_param_grads_workspace_bytes(rows: int, D: int) -> int:
    return _lib.load().vitae_layernorm_param_grads_workspace_bytes(rows, D)


def layernorm_param_grads(dy, x, mean, rstd, dx_out, workspace, dgamma=None, dbeta=None, dbias=None,
                          accumulate: bool = False, dy2=None) -> None:
    """dgamma = sum dy*xhat, dbeta = sum dy, dbias = column sums of dx_out, one launch; workspace: zero-filled uint8 tensor of
    >= layernorm_param_grads_workspace_bytes(rows, D) bytes (see header)."""
    lib = _lib.load()
    rows, D = x.numel() // x.shape[-1], x.shape[-1]
    if workspace.numel() * workspace.element_size() < lib.vitae_layernorm_param_grads_workspace_bytes(rows, D):
        raise _lib.VitaeError("layernorm_param_grads: workspace too small")
    dy16, dy32 = _dy_pair(dy, dy2)
    check(lib.vitae_layernorm_param_grads(dy16, dy32, x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _ptr(dx_out),
                                          workspace.data_ptr(), _ptr(dgamma), _ptr(dbeta), _ptr(dbias), int(accumulate), rows,
                                          D, _stream()), "vitae_layernorm_param_grads")


def col_job(a, out0, *, x=None, mean=None, rstd=None, a2=None, out1=None, cols: Optional[int] = None,
            ld: Optional[int] = None):
    """One job of block_colreduce: plain column sum of ``a`` [rows, cols] (bf16 / fp32) into out0, or, with x / mean / rstd,
    the LayerNorm affine gradients (out0 = dgamma, out1 = dbeta) for the upstream gradient a (+ a2)."""
    cols = a.shape[-1] if cols is None else cols
    return _lib.ColJob(a.data_ptr(), _ptr(a2), _ptr(x), _ptr(mean), _ptr(rstd), out0.data_ptr(), _ptr(out1), cols,
                       cols if ld is None else ld, int(a.dtype == _BF16), 0)


def block_colreduce_workspace_bytes(rows: int, cols_list) -> int:
    return _lib.load().vitae_block_colreduce_workspace_bytes(rows, sum((c + 255) // 256 * 256 for c in cols_list))


def block_colreduce(jobs, rows: int, workspace, accumulate: bool = False) -> None:
    """All column reductions of one block's backward in one launch (include/vitae_b200.h)."""
    lib = _lib.load()
    arr = (_lib.ColJob * len(jobs))(*jobs)
    check(lib.vitae_block_colreduce(arr, len(jobs), rows, int(accumulate), workspace.data_ptr(),
                                    workspace.numel() * workspace.element_size(), _stream()), "vitae_block_colreduce")


def colsum(inp, rows: int, cols: int, out, workspace, accumulate: bool = False, ld: Optional[int] = None,
           scale_ptr: Optional[torch.Tensor] = None) -> None:
    """workspace: uint8 tensor of >= colsum_workspace_bytes(rows, cols) bytes, zero-filled at allocation (see header)."""
    lib = _lib.load()
    if workspace.numel() * workspace.element_size() < lib.vitae_colsum_workspace_bytes(rows, cols):
        raise _lib.VitaeError("colsum: workspace too small")
    in16 = inp.data_ptr() if inp.dtype == _BF16 else None
    in32 = inp.data_ptr() if inp.dtype == _F32 else None
    check(lib.vitae_colsum(in16, in32, rows, cols, ld if ld is not None else cols, out.data_ptr(), int(accumulate),
                           workspace.data_ptr(), _ptr(scale_ptr), _stream()), "vitae_colsum")


def attention_fwd(qkv, out, lse, B: int, N: int, H: int, hd: int, scale: float) -> None:
    lib = _lib.load()
    _req(qkv, _BF16, "attention qkv")
    check(lib.vitae_attention_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, hd, scale, _stream()),
          "vitae_attention_fwd")


def attention_bwd(qkv, out, dout, lse, delta, dqkv, B: int, N: int, H: int, hd: int, scale: float) -> None:
    lib = _lib.load()
    check(lib.vitae_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                  dqkv.data_ptr(), B, N, H, hd, scale, _stream()), "vitae_attention_bwd")


def random_masking(noise, ids_shuffle, ids_restore, mask, len_keep: int) -> None:
    lib = _lib.load()
    _req(noise, _F32, "noise")
    B, L = noise.shape
    check(lib.vitae_random_masking(noise.data_ptr(), ids_shuffle.data_ptr(), ids_restore.data_ptr(), mask.data_ptr(),
                                   B, L, len_keep, _stream()), "vitae_random_masking")


def build_row_maps(ids_shuffle, keep: int, out: Optional[dict] = None) -> dict:
    """Row maps of include/vitae_b200.h::vitae_build_row_maps; allocates them unless ``out`` holds the buffers."""
    lib = _lib.load()
    _req(ids_shuffle, _I32, "ids_shuffle")
    B, L = ids_shuffle.shape
    Ne = keep + 1
    if out is None:
        dev = ids_shuffle.device
        sizes = {"enc_tok_rows": B * keep, "enc_cls_rows": B, "pe_pos_rows": B * keep, "dec_rows_of_enc": B * Ne,
                 "dec_pos_rows_of_enc": B * Ne, "masked_dec_rows": max(1, B * (L - keep)),
                 "masked_pos_rows": max(1, B * (L - keep))}
        out = {k: torch.empty(n, dtype=_I32, device=dev) for k, n in sizes.items()}
    check(lib.vitae_build_row_maps(ids_shuffle.data_ptr(), B, L, keep, out["enc_tok_rows"].data_ptr(),
                                   out["enc_cls_rows"].data_ptr(), out["pe_pos_rows"].data_ptr(),
                                   out["dec_rows_of_enc"].data_ptr(), out["dec_pos_rows_of_enc"].data_ptr(),
                                   out["masked_dec_rows"].data_ptr(), out["masked_pos_rows"].data_ptr(), _stream()),
          "vitae_build_row_maps")
    return out


def im2col_patches(vol, ids_shuffle, cols, p: int, keep: int) -> None:
    lib = _lib.load()
    _req(vol, _F32, "volume")
    B, C, V = vol.shape[0], vol.shape[1], vol.shape[2]
    L = (V // p) ** 3
    check(lib.vitae_im2col_patches(vol.data_ptr(), ids_shuffle.data_ptr(), cols.data_ptr(), B, C, V, p, L, keep,
                                   _stream()), "vitae_im2col_patches")


def fill_rows(dst, row_idx, nrows: int, D: int, src0, src0_rows=None, src1=None, src1_rows=None) -> None:
    lib = _lib.load()
    check(lib.vitae_fill_rows(dst.data_ptr(), _ptr(row_idx), nrows, D, src0.data_ptr(), _ptr(src0_rows), _ptr(src1),
                              _ptr(src1_rows), _stream()), "vitae_fill_rows")


def gather_rows(src, row_idx, nrows: int, D: int, dst_bf16=None, dst_f32=None) -> None:
    lib = _lib.load()
    check(lib.vitae_gather_rows(src.data_ptr(), _ptr(row_idx), nrows, D, _ptr(dst_bf16), _ptr(dst_f32), _stream()),
          "vitae_gather_rows")


def sum_rows(src, row_idx, nrows: int, D: int, out, accumulate: bool = False) -> None:
    lib = _lib.load()
    check(lib.vitae_sum_rows(src.data_ptr(), _ptr(row_idx), nrows, D, out.data_ptr(), int(accumulate), _stream()),
          "vitae_sum_rows")


def masked_mse_fwd(pred, vol, mask, patch_sums, loss_out, p: int) -> None:
    lib = _lib.load()
    B, C, V = vol.shape[0], vol.shape[1], vol.shape[2]
    check(lib.vitae_masked_mse_fwd(pred.data_ptr(), int(pred.dtype == _BF16), vol.data_ptr(), mask.data_ptr(),
                                   patch_sums.data_ptr(), loss_out.data_ptr(), B, C, V, p, _stream()),
          "vitae_masked_mse_fwd")


def masked_mse_bwd(pred, vol, mask, mask_sum, dloss, dpred, p: int) -> None:
    lib = _lib.load()
    B, C, V = vol.shape[0], vol.shape[1], vol.shape[2]
    check(lib.vitae_masked_mse_bwd(pred.data_ptr(), int(pred.dtype == _BF16), vol.data_ptr(), mask.data_ptr(),
                                   mask_sum.data_ptr(), dloss.data_ptr(), dpred.data_ptr(), B, C, V, p, _stream()),
          "vitae_masked_mse_bwd")


def edge_scratch_floats(B: int, C: int, V: int) -> int:
    return _lib.load().vitae_edge_scratch_floats(B, C, V)


def gaussian_taps(sigma: float = 2.0):
    """The reference's 1-D taps (model/model_utils/gaussian_filter.py:5-13, incl. the linspace(-ks//2, ks//2+1, ks) quirk)
    as a python list of fp32 values."""
    ks = int(sigma * 5)
    ks += 1 - ks % 2
    ts = torch.linspace(-ks // 2, ks // 2 + 1, ks)
    t = torch.exp(-(ts / sigma) ** 2 / 2)
    return (t / t.sum()).tolist()


def edge_target(vol, taps, scratch, e_tgt) -> None:
    lib = _lib.load()
    _req(vol, _F32, "edge volume")
    B, C, V = vol.shape[0], vol.shape[1], vol.shape[2]
    host = (ctypes.c_float * len(taps))(*taps)
    check(lib.vitae_edge_target(vol.data_ptr(), ctypes.cast(host, ctypes.c_void_p), len(taps), scratch.data_ptr(),
                                e_tgt.data_ptr(), B, C, V, _stream()), "vitae_edge_target")


def edge_loss_fwd(pred_bf16, e_tgt, scratch, resid, loss_out, B: int, C: int, V: int, p: int) -> None:
    lib = _lib.load()
    _req(pred_bf16, _BF16, "edge pred")
    check(lib.vitae_edge_loss_fwd(pred_bf16.data_ptr(), e_tgt.data_ptr(), scratch.data_ptr(), resid.data_ptr(),
                                  loss_out.data_ptr(), B, C, V, p, _stream()), "vitae_edge_loss_fwd")


def edge_loss_bwd(resid, scratch, upstream, dpred_bf16, B: int, C: int, V: int, p: int) -> None:
    lib = _lib.load()
    check(lib.vitae_edge_loss_bwd(resid.data_ptr(), scratch.data_ptr(), upstream.data_ptr(), dpred_bf16.data_ptr(), B, C, V,
                                  p, _stream()), "vitae_edge_loss_bwd")


def prefetch_l2(tensors) -> None:
    """Pulls the storage of up to 12 tensors (contiguous) into L2 on the current stream (include/vitae_b200.h)."""
    ts = [t for t in tensors if t is not None and t.numel() > 0]
    if not ts:
        return
    lib = _lib.load()
    for i in range(0, len(ts), 12):
        part = ts[i:i + 12]
        ptrs = (ctypes.c_void_p * len(part))(*[t.data_ptr() for t in part])
        sizes = (ctypes.c_size_t * len(part))(*[t.numel() * t.element_size() for t in part])
        check(lib.vitae_prefetch_l2(ptrs, sizes, len(part), _stream()), "vitae_prefetch_l2")


def cast_params_bf16(table, ntensors: int, dst, total: int) -> None:
    lib = _lib.load()
    check(lib.vitae_cast_params_bf16(table.data_ptr(), ntensors, dst.data_ptr(), total, _stream()),
          "vitae_cast_params_bf16")


def adamw_step(param, grad, exp_avg, exp_avg_sq, param_bf16, n: int, lr: float, beta1: float, beta2: float,
               eps: float, weight_decay: float, step: int, inv_scale=None, found_inf=None) -> None:
    lib = _lib.load()
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    check(lib.vitae_adamw_step(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                               _ptr(param_bf16), n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, _ptr(inv_scale),
                               _ptr(found_inf), _stream()), "vitae_adamw_step")


def optim_workspace_bytes() -> int:
    return _lib.load().vitae_optim_workspace_bytes()


def optim_prepare(grad, n: int, ctl, workspace, growth_factor: float, backoff_factor: float, growth_interval: int,
                  use_scaler: bool, grad2=None, n2: int = 0) -> None:
    lib = _lib.load()
    _req(grad, _F32, "optim grad"); _req(ctl, _F32, "optim ctl")
    check(lib.vitae_optim_prepare(grad.data_ptr(), n, _ptr(grad2), n2, ctl.data_ptr(), workspace.data_ptr(), growth_factor,
                                  backoff_factor, growth_interval, int(use_scaler), _stream()), "vitae_optim_prepare")


def adamw_flat(param, grad, exp_avg, exp_avg_sq, param_bf16, n: int, group_of_chunk, hyper_rows, ctl, start: int = 0,
               max_blocks: int = 0) -> None:
    """AdamW over elements [start, start + n) of the flat buffers (start, n multiples of 64).  hyper_rows: python list of
    (lr, beta1, beta2, eps, weight_decay) per parameter group (passed by value)."""
    lib = _lib.load()
    ng = len(hyper_rows)
    host = (ctypes.c_float * (8 * ng))()
    for i, row in enumerate(hyper_rows):
        for j, val in enumerate(row):
            host[8 * i + j] = float(val)
    assert start % 64 == 0 and n % 64 == 0
    p16 = None if param_bf16 is None else param_bf16.data_ptr() + 2 * start
    check(lib.vitae_adamw_flat(param.data_ptr() + 4 * start, grad.data_ptr() + 4 * start, exp_avg.data_ptr() + 4 * start,
                               exp_avg_sq.data_ptr() + 4 * start, p16, n, group_of_chunk.data_ptr() + start // 64,
                               ctypes.cast(host, ctypes.c_void_p), ng, ctl.data_ptr(), max_blocks, _stream()),
          "vitae_adamw_flat")


def peer_table(ptrs) -> "ctypes.Array":
    """Host array of the ranks' base addresses of one symmetric buffer (index = rank), as the dp_* entry points take it."""
    arr = (ctypes.c_void_p * 8)()
    for i, p in enumerate(ptrs):
        arr[i] = int(p)
    return arr


def dp_owned_elems(lo: int, hi: int, granule_shift: int, world: int, rank: int) -> int:
    return int(_lib.load().vitae_dp_owned_elems(lo, hi, granule_shift, world, rank))


def dp_reduce_shard_blocks(lo: int, hi: int, granule_shift: int, world: int, rank: int, max_blocks: int = 0) -> int:
    return int(_lib.load().vitae_dp_reduce_shard_blocks(lo, hi, granule_shift, world, rank, max_blocks))


def dp_reduce_shard(grad_peers, world: int, rank: int, lo: int, hi: int, granule_shift: int, inv_world: float, partials,
                    max_blocks: int = 0) -> int:
    """This rank's part (granules q with q % world == rank) of flat gradient elements [lo, hi) := mean over ranks (peer
    copies read over NVLink, ``grad_peers`` = peer_table of the gradient buffers); per-block sums of squares of the result
    -> ``partials`` (fp32 view).  Returns the number of partials written (0: the rank owns nothing of the slice)."""
    _req(partials, _F32, "partials")
    lib = _lib.load()
    nb = int(lib.vitae_dp_reduce_shard_blocks(lo, hi, granule_shift, world, rank, max_blocks))
    assert partials.numel() >= nb
    check(lib.vitae_dp_reduce_shard(grad_peers, world, rank, lo, hi, granule_shift, float(inv_world), partials.data_ptr(),
                                    max_blocks, _stream()), "vitae_dp_reduce_shard")
    return nb


def adamw_shard(param_peers, param_bf16_peers, world: int, rank: int, lo: int, hi: int, granule_shift: int, grad, exp_avg,
                exp_avg_sq, group_of_chunk, f32_chunk, hyper_rows, ctl, max_blocks: int = 0) -> None:
    """adamw_flat over this rank's part of [lo, hi) (grad / exp_avg / exp_avg_sq / group_of_chunk / f32_chunk: whole flat
    buffers), with the all-gather of the result fused in: bf16 shadow to every rank, fp32 master to the peers for the
    chunks flagged in ``f32_chunk`` (uint8 per 64 elements)."""
    lib = _lib.load()
    ng = len(hyper_rows)
    host = (ctypes.c_float * (8 * ng))()
    for i, row in enumerate(hyper_rows):
        for j, val in enumerate(row):
            host[8 * i + j] = float(val)
    check(lib.vitae_adamw_shard(param_peers, param_bf16_peers, world, rank, lo, hi, granule_shift, grad.data_ptr(),
                                exp_avg.data_ptr(), exp_avg_sq.data_ptr(), group_of_chunk.data_ptr(), f32_chunk.data_ptr(),
                                ctypes.cast(host, ctypes.c_void_p), ng, ctl.data_ptr(), max_blocks, _stream()),
          "vitae_adamw_shard")


def sum_partials(partials, n: int, out) -> None:
    _req(partials, _F32, "partials"); _req(out, _F32, "out")
    check(_lib.load().vitae_sum_partials(partials.data_ptr(), n, out.data_ptr(), _stream()), "vitae_sum_partials")


def optim_finalize_peers(partial_peers, world: int, count: int, ctl, growth_factor: float, backoff_factor: float,
                         growth_interval: int, use_scaler: bool) -> None:
    """optim_finalize over ``count`` partial sums of squares from EVERY rank (peer_table of the ranks' partial buffers,
    summed rank-major in double: every rank computes the same control block)."""
    check(_lib.load().vitae_optim_finalize_peers(partial_peers, world, count, ctl.data_ptr(), float(growth_factor),
                                                 float(backoff_factor), int(growth_interval), int(bool(use_scaler)), _stream()),
          "vitae_optim_finalize_peers")


def cast_f32_to_bf16(src, dst, max_blocks: int = 0) -> None:
    _req(src, _F32, "cast src"); _req(dst, _BF16, "cast dst")
    check(_lib.load().vitae_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), max_blocks, _stream()),
          "vitae_cast_f32_to_bf16")


def cast_bf16_to_f32(src, dst, max_blocks: int = 0) -> None:
    _req(src, _BF16, "cast src"); _req(dst, _F32, "cast dst")
    check(_lib.load().vitae_cast_bf16_to_f32(src.data_ptr(), dst.data_ptr(), src.numel(), max_blocks, _stream()),
          "vitae_cast_bf16_to_f32")


RAW_TYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.uint16: 3, torch.int16: 4, torch.uint8: 5}
INGEST_MODES = {"z_score_channel": 0, "z_score_sample": 1, "min_max": 2}


def ingest_workspace_bytes(B: int, C: int) -> int:
    return _lib.load().vitae_ingest_workspace_bytes(B, C)


def ingest_normalize(raw, out, mode: str, workspace, stats=None) -> None:
    """raw [B, C, ...] in its storage dtype (RAW_TYPES) -> out fp32, normalised per ``mode`` (INGEST_MODES; the reference's
    Dataset._normalize_data, dataset/egd_dataset/egd.py:44-50 / dataset/brats_dataset/brats.py:26-32)."""
    if raw.dtype not in RAW_TYPES:
        raise _lib.VitaeError(f"ingest_normalize: unsupported raw dtype {raw.dtype}")
    if mode not in INGEST_MODES:
        raise _lib.VitaeError(f"ingest_normalize: mode {mode!r} not in {sorted(INGEST_MODES)}")
    if not (raw.is_cuda and raw.is_contiguous()):
        raise _lib.VitaeError("ingest_normalize: raw volume must be a contiguous CUDA tensor")
    _req(out, _F32, "ingest out")
    B, C = raw.shape[0], raw.shape[1]
    vox = raw[0, 0].numel()
    check(_lib.load().vitae_ingest_normalize(raw.data_ptr(), RAW_TYPES[raw.dtype], out.data_ptr(), B, C, vox,
                                             INGEST_MODES[mode], workspace.data_ptr(), _ptr(stats), _stream()),
          "vitae_ingest_normalize")


def bn_relu_fwd(h, gamma, beta, eps: float, act_bf16, mean, rstd, running_mean=None, running_var=None,
                momentum: float = 0.1) -> None:
    _req(h, _F32, "bn h"); _req(act_bf16, _BF16, "bn act")
    M, D = h.shape
    check(_lib.load().vitae_bn_relu_fwd(h.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, act_bf16.data_ptr(), mean.data_ptr(),
                                        rstd.data_ptr(), _ptr(running_mean), _ptr(running_var), momentum, M, D, _stream()),
          "vitae_bn_relu_fwd")


def bn_relu_bwd(dact_bf16, h, gamma, beta, mean, rstd, dh_bf16, dgamma, dbeta, accumulate: bool = False) -> None:
    _req(dact_bf16, _BF16, "bn dact"); _req(dh_bf16, _BF16, "bn dh")
    M, D = h.shape
    check(_lib.load().vitae_bn_relu_bwd(dact_bf16.data_ptr(), h.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(),
                                        rstd.data_ptr(), dh_bf16.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), int(accumulate),
                                        M, D, _stream()), "vitae_bn_relu_bwd")


def cosine_loss_workspace_floats(M: int) -> int:
    return _lib.load().vitae_cosine_loss_workspace_floats(M)


def cosine_loss_fwd(p1, z2, p2, z1, weight: float, workspace, loss) -> None:
    for t in (p1, z2, p2, z1):
        _req(t, _F32, "cosine operand")
    M, D = p1.shape
    check(_lib.load().vitae_cosine_loss_fwd(p1.data_ptr(), z2.data_ptr(), p2.data_ptr(), z1.data_ptr(), M, D, weight,
                                            workspace.data_ptr(), loss.data_ptr(), _stream()), "vitae_cosine_loss_fwd")


def cosine_loss_bwd(p1, z2, p2, z1, weight: float, workspace, upstream, dp1, dp2) -> None:
    M, D = p1.shape
    check(_lib.load().vitae_cosine_loss_bwd(p1.data_ptr(), z2.data_ptr(), p2.data_ptr(), z1.data_ptr(), M, D, weight,
                                            workspace.data_ptr(), upstream.data_ptr(), dp1.data_ptr(), dp2.data_ptr(), _stream()),
          "vitae_cosine_loss_bwd")


def grad_sqnorm_blocks(n: int, max_blocks: int = 0) -> int:
    return _lib.load().vitae_grad_sqnorm_blocks(n, max_blocks)


def grad_sqnorm(grad, partials, max_blocks: int = 0) -> None:
    """Partial sums of squares of a flat fp32 gradient slice -> partials[: grad_sqnorm_blocks(n, max_blocks)]."""
    _req(grad, _F32, "sqnorm grad"); _req(partials, _F32, "sqnorm partials")
    check(_lib.load().vitae_grad_sqnorm(grad.data_ptr(), grad.numel(), partials.data_ptr(), max_blocks, _stream()),
          "vitae_grad_sqnorm")


def optim_finalize(partials, npartials: int, ctl, growth_factor: float, backoff_factor: float, growth_interval: int,
                   use_scaler: bool) -> None:
    check(_lib.load().vitae_optim_finalize(partials.data_ptr(), npartials, ctl.data_ptr(), growth_factor, backoff_factor,
                                           growth_interval, int(use_scaler), _stream()), "vitae_optim_finalize")


def pred_mse_partial_floats(M: int, P: int, tile_n: int = 128) -> int:
    return _lib.load().vitae_pred_mse_partial_floats(M, P, tile_n)


def gemm_pred_mse(hN, W, bias, B: int, L: int, Dd: int, vol, mask, p: int, mask_sum: float, pred_bf16, g_bf16, partials,
                  tile_n: int = 128) -> None:
    """decoder_pred GEMM with the masked reconstruction loss in its epilogue (include/vitae_b200.h: vitae_gemm_pred_mse)."""
    _req(hN, _BF16, "pred_mse hN"); _req(W, _BF16, "pred_mse W"); _req(vol, _F32, "pred_mse vol"); _req(mask, _F32, "pred_mse mask")
    _req(pred_bf16, _BF16, "pred_mse pred"); _req(g_bf16, _BF16, "pred_mse g"); _req(partials, _F32, "pred_mse partials")
    C, V = vol.shape[1], vol.shape[2]
    check(_lib.load().vitae_gemm_pred_mse(hN.data_ptr(), W.data_ptr(), _ptr(bias), B, L, Dd, vol.data_ptr(), mask.data_ptr(), C, V, p,
                                          float(mask_sum), pred_bf16.data_ptr(), g_bf16.data_ptr(), partials.data_ptr(), tile_n,
                                          _stream()), "vitae_gemm_pred_mse")
    if _gemm_log is not None:
        M, P = B * (L + 1), p ** 3 * C
        _gemm_log.append((2.0 * M * P * Dd, lambda: gemm_pred_mse(hN, W, bias, B, L, Dd, vol, mask, p, mask_sum, pred_bf16, g_bf16,
                                                                  partials, tile_n)))


def pred_mse_finalize(partials, P: int, mask_sum: float, loss_out) -> None:
    check(_lib.load().vitae_pred_mse_finalize(partials.data_ptr(), partials.numel(), P, float(mask_sum), loss_out.data_ptr(),
                                              _stream()), "vitae_pred_mse_finalize")
