#!/usr/bin/env python
"""Per-phase timeline of the tcgen05 GEMM kernel (debug build: VITAE_TRACE=1 python -m vit_ae_plus_plus_b200.build --force).
Each CTA stamps %globaltimer (ns) at: 0 entry, 1 prologue done (barriers, TMEM), 2 PDL wait passed, 3 first TMA issued,
4 last TMA issued, 5 first stage landed, 6 last MMA committed, 7 accumulator ready (epilogue starts), 8 epilogue done,
9 CTA done.  Prints, per shape, the median over CTAs of each phase relative to the earliest CTA entry."""
import ctypes
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vit_ae_plus_plus_b200 import _lib, ops  # noqa: E402
from tools.ncu_gemm import SHAPES  # noqa: E402

NAMES = ["entry", "prologue", "pdl_wait", "tma_first", "tma_last", "stage0_landed", "mma_commit", "acc_ready", "epi_done",
         "cta_done", "chunk0_staged", "chunk0_stored"]


def main():
    dev = torch.device("cuda")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    ws = ops.GrowBuf(dev)
    trace = torch.zeros(1 << 16, 16, dtype=torch.int64, device=dev)
    lib.vitae_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
    assert lib.vitae_debug_set_gemm_trace(trace.data_ptr()) == 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, M, N, K, amn, bmn, out, (tn, sk) in SHAPES:
        A = torch.randn((K, M) if amn else (M, K), device=dev).bfloat16()
        B = torch.randn((K, N) if bmn else (N, K), device=dev).bfloat16()
        o = torch.empty(M, N, device=dev, dtype=torch.float32 if out == "f32" else torch.bfloat16)
        kw = {"out_f32": o} if out == "f32" else {"out_bf16": o}
        for mode in ("warm", "cold"):
            for rep in range(3):
                if mode == "cold":
                    flush.zero_()
                trace.zero_()
                torch.cuda.synchronize()
                ops.gemm(A, B, M, N, K, a_mn_major=bool(amn), b_mn_major=bool(bmn), workspace=ws, tile_n=tn, split_k=sk, **kw)
                torch.cuda.synchronize()
            t = trace.cpu()
            live = t[:, 0] > 0
            t = t[live]
            t0 = int(t[:, 0].min())
            cols = []
            for k in range(12):
                v = [int(x) - t0 for x in t[:, k].tolist() if x > 0]
                cols.append((statistics.median(v), max(v)) if v else (0, 0))
            print(f"{name:14s} {mode} tile_n={tn} split={sk} ctas={t.shape[0]:4d}  " +
                  "  ".join(f"{n}={m / 1000:.2f}/{mx / 1000:.2f}" for n, (m, mx) in zip(NAMES, cols)) + "  (us: median/max)")


if __name__ == "__main__":
    main()
