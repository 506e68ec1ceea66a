#!/usr/bin/env python
"""Times every distinct GEMM shape of one training step in isolation (CUDA-graph of REPS back-to-back launches, CUDA
events) for each tile_n / split_k candidate.  Output: one line per (shape, config) with us per launch and TFLOP/s.
usage (GPU box): python tools/gemm_bench.py [--batch 4] > gpurun_out/gemm_bench.txt"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vit_ae_plus_plus_b200 import ops  # noqa: E402

REPS = 20


def shapes(B, keep=128, L=512, D=768, hid=3072, Dd=512, hidd=2048, P=16384):
    Me, Md = B * (keep + 1), B * (L + 1)
    out = []
    for tag, M, d, h in (("enc", Me, D, hid), ("dec", Md, Dd, hidd)):
        out += [(f"{tag}.qkv.fwd", M, 3 * d, d, 0, 0), (f"{tag}.proj.fwd", M, d, d, 0, 0), (f"{tag}.fc1.fwd", M, h, d, 0, 0),
                (f"{tag}.fc2.fwd", M, d, h, 0, 0),
                (f"{tag}.fc2.dgrad", M, h, d, 0, 1), (f"{tag}.fc1.dgrad", M, d, h, 0, 1), (f"{tag}.proj.dgrad", M, d, d, 0, 1),
                (f"{tag}.qkv.dgrad", M, d, 3 * d, 0, 1),
                (f"{tag}.fc2.wgrad", d, h, M, 1, 1), (f"{tag}.fc1.wgrad", h, d, M, 1, 1), (f"{tag}.proj.wgrad", d, d, M, 1, 1),
                (f"{tag}.qkv.wgrad", 3 * d, d, M, 1, 1)]
    out += [("patch_embed.fwd", B * keep, D, P, 0, 0), ("patch_embed.wgrad", D, P, B * keep, 1, 1),
            ("dec_embed.fwd", Me, Dd, D, 0, 0), ("dec_embed.dgrad", Me, D, Dd, 0, 1), ("dec_embed.wgrad", Dd, D, Me, 1, 1),
            ("pred.fwd", Md, P, Dd, 0, 0), ("pred.dgrad", Md, Dd, P, 0, 1), ("pred.wgrad", P, Dd, Md, 1, 1)]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    a = ap.parse_args()
    dev = torch.device("cuda")
    ws = ops.GrowBuf(dev)
    print(f"# batch {a.batch}; us per launch inside a graph of {REPS} back-to-back launches")
    for name, M, N, K, amn, bmn in shapes(a.batch):
        A = torch.randn((K, M) if amn else (M, K), device=dev).bfloat16()
        Bm = torch.randn((K, N) if bmn else (N, K), device=dev).bfloat16()
        out32 = torch.empty(M, N, device=dev)
        out16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        auto = ops.gemm_config(M, N, K)
        cands = []
        for tn in (64, 128, 256):
            for sk in (1, 2, 4, 8):
                if sk > 1 and (K + 63) // 64 < 4 * sk:
                    continue
                if N % 64 and tn == 64:
                    continue
                cands.append((tn, sk))
        best = None
        for tn, sk in cands:
            kw = dict(a_mn_major=bool(amn), b_mn_major=bool(bmn), tile_n=tn, split_k=sk, workspace=ws)
            if "wgrad" in name:
                kw["out_f32"] = out32
            else:
                kw["out_bf16"] = out16
            try:
                ops.gemm(A, Bm, M, N, K, **kw)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(REPS):
                        ops.gemm(A, Bm, M, N, K, **kw)
                g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / (5 * REPS)
            except Exception as ex:  # noqa: BLE001
                print(f"{name:20s} M={M:6d} N={N:6d} K={K:6d} tile_n={tn:3d} split={sk} FAILED {ex}")
                continue
            tf = 2.0 * M * N * K / us / 1e6
            mark = " <auto" if (tn, sk) == auto else ""
            print(f"{name:20s} M={M:6d} N={N:6d} K={K:6d} tile_n={tn:3d} split={sk} {us:8.2f} us {tf:7.1f} TF/s{mark}", flush=True)
            if best is None or us < best[0]:
                best = (us, tn, sk)
        print(f"  -> best {name}: tile_n={best[1]} split={best[2]} {best[0]:.2f} us")


if __name__ == "__main__":
    main()
