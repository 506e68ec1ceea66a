"""ctypes binding of libvitae_b200.so (C ABI declared in include/vitae_b200.h).

The product path has NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_longlong, c_size_t, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VITAE_LIB") or os.path.join(PKG_DIR, "libvitae_b200.so")   # VITAE_LIB: an experiment build


class VitaeError(RuntimeError):
    pass


class GemmEpilogue(Structure):
    # mirrors struct vitae_gemm_epilogue (include/vitae_b200.h)
    _fields_ = [
        ("alpha", c_float),
        ("alpha_ptr", c_void_p),
        ("bias", c_void_p),
        ("addend", c_void_p),
        ("add_rows", c_void_p),
        ("ldadd", c_int32),
        ("dgelu_src", c_void_p),
        ("ld_dgelu", c_int32),
        ("out_f32", c_void_p),
        ("ld_f32", c_int32),
        ("accumulate", c_int32),
        ("out_bf16", c_void_p),
        ("out_gelu_bf16", c_void_p),
        ("ld_bf16", c_int32),
        ("out_rows", c_void_p),
    ]


class ColJob(Structure):
    # mirrors struct vitae_col_job (include/vitae_b200.h)
    _fields_ = [("a", c_void_p), ("a2", c_void_p), ("x", c_void_p), ("mean", c_void_p), ("rstd", c_void_p),
                ("out0", c_void_p), ("out1", c_void_p), ("cols", c_int32), ("ld", c_int32), ("a_is_bf16", c_int32),
                ("reserved", c_int32)]


# name -> (restype, argtypes); every symbol declared in include/vitae_b200.h
SIGNATURES = {
    "vitae_abi_version": (c_int, []),
    "vitae_last_error": (c_char_p, []),
    "vitae_check_device": (c_int, []),
    "vitae_launch_count": (c_longlong, []),
    "vitae_gemm_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                POINTER(GemmEpilogue), c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "vitae_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "vitae_gemm_workspace_bytes_for": (c_size_t, [POINTER(GemmEpilogue), c_int, c_int, c_int, c_int, c_int]),
    "vitae_gemm_autotune": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    POINTER(GemmEpilogue), c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, POINTER(c_int),
                                    POINTER(c_int)]),
    "vitae_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                    c_float, c_void_p]),
    "vitae_layernorm_bwd": (c_int, [c_void_p] * 9 + [c_int, c_int, c_void_p]),
    "vitae_layernorm_param_grads": (c_int, [c_void_p] * 10 + [c_int, c_int, c_int, c_void_p]),
    "vitae_layernorm_param_grads_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vitae_layernorm_bwd_blocks": (c_int, [c_int]),
    "vitae_reduce_partials": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "vitae_block_colreduce": (c_int, [POINTER(ColJob), c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "vitae_block_colreduce_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vitae_colsum": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "vitae_pred_mse_partial_floats": (c_size_t, [c_int, c_int, c_int]),
    "vitae_gemm_pred_mse": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_float, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "vitae_pred_mse_finalize": (c_int, [c_void_p, c_longlong, c_int, c_float, c_void_p, c_void_p]),
    "vitae_colsum_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vitae_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "vitae_attention_bwd": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "vitae_random_masking": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vitae_build_row_maps": (c_int, [c_void_p, c_int, c_int, c_int] + [c_void_p] * 8),
    "vitae_im2col_patches": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "vitae_fill_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vitae_gather_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vitae_sum_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "vitae_masked_mse_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p]),
    "vitae_masked_mse_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                     c_int, c_int, c_void_p]),
    "vitae_edge_scratch_floats": (c_size_t, [c_int, c_int, c_int]),
    "vitae_edge_target": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vitae_edge_loss_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vitae_edge_loss_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vitae_prefetch_l2": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "vitae_cast_params_bf16": (c_int, [c_void_p, c_int, c_void_p, c_longlong, c_void_p]),
    "vitae_adamw_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_float, c_float, c_float,
                                 c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "vitae_optim_prepare": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p, c_float, c_float, c_int,
                                    c_int, c_void_p]),
    "vitae_optim_workspace_bytes": (c_size_t, []),
    "vitae_grad_sqnorm_blocks": (c_int, [c_longlong, c_int]),
    "vitae_grad_sqnorm": (c_int, [c_void_p, c_longlong, c_void_p, c_int, c_void_p]),
    "vitae_optim_finalize": (c_int, [c_void_p, c_int, c_void_p, c_float, c_float, c_int, c_int, c_void_p]),
    "vitae_adamw_flat": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_int,
                                 c_void_p, c_int, c_void_p]),
    "vitae_dp_owned_elems": (c_longlong, [c_longlong, c_longlong, c_int, c_int, c_int]),
    "vitae_dp_reduce_shard_blocks": (c_int, [c_longlong, c_longlong, c_int, c_int, c_int, c_int]),
    "vitae_dp_reduce_shard": (c_int, [c_void_p, c_int, c_int, c_longlong, c_longlong, c_int, c_float, c_void_p, c_int, c_void_p]),
    "vitae_adamw_shard": (c_int, [c_void_p, c_void_p, c_int, c_int, c_longlong, c_longlong, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "vitae_sum_partials": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "vitae_optim_finalize_peers": (c_int, [c_void_p, c_int, c_int, c_void_p, c_float, c_float, c_int, c_int, c_void_p]),
    "vitae_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "vitae_cast_bf16_to_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "vitae_bn_relu_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                  c_int, c_int, c_void_p]),
    "vitae_bn_relu_bwd": (c_int, [c_void_p] * 9 + [c_int, c_int, c_int, c_void_p]),
    "vitae_cosine_loss_workspace_floats": (c_size_t, [c_int]),
    "vitae_cosine_loss_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "vitae_cosine_loss_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "vitae_ingest_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vitae_ingest_normalize": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_longlong, c_int, c_void_p, c_void_p,
                                       c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Loads the library (once).  Raises VitaeError when it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VitaeError(
            f"{LIB_PATH} not found: build it with `python -m vit_ae_plus_plus_b200.build` "
            "(nvcc, sm_100a).  This package has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.vitae_abi_version() != 1:
        raise VitaeError("libvitae_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vitae_last_error().decode(errors="replace")
        raise VitaeError(f"{what} failed (rc={rc}): {msg}")
