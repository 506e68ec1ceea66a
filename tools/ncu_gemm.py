#!/usr/bin/env python
"""Launches a few representative GEMM shapes of the training step a handful of times each, for `ncu --set full`.
usage (GPU box): ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 2 -c 12 \
                     -o gpurun_out/gemm_full python tools/ncu_gemm.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vit_ae_plus_plus_b200 import ops  # noqa: E402

SHAPES = [  # name, M, N, K, a_mn, b_mn, out, (tile_n, split_k) as picked by the on-device autotuner (profiles/)
    ("dec.proj.fwd", 2052, 512, 512, 0, 0, "bf16", (64, 1)),
    ("enc.qkv.fwd", 516, 2304, 768, 0, 0, "bf16", (128, 1)),
    ("dec.fc1.fwd", 2052, 2048, 512, 0, 0, "bf16", (128, 1)),
    ("enc.fc2.wgrad", 768, 3072, 516, 1, 1, "f32", (128, 1)),
    ("pred.fwd", 2052, 16384, 512, 0, 0, "bf16", (128, 1)),
    ("pred.wgrad", 16384, 512, 2052, 1, 1, "f32", (128, 1)),
]


def main():
    dev = torch.device("cuda")
    ws = ops.GrowBuf(dev)
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    for name, M, N, K, amn, bmn, out, (tn, sk) in SHAPES:
        A = torch.randn((K, M) if amn else (M, K), device=dev).bfloat16()
        B = torch.randn((K, N) if bmn else (N, K), device=dev).bfloat16()
        o = torch.empty(M, N, device=dev, dtype=torch.float32 if out == "f32" else torch.bfloat16)
        kw = {"out_f32": o} if out == "f32" else {"out_bf16": o}
        for _ in range(reps):
            ops.gemm(A, B, M, N, K, a_mn_major=bool(amn), b_mn_major=bool(bmn), workspace=ws, tile_n=tn, split_k=sk, **kw)
        torch.cuda.synchronize()
        print(name, tn, sk)


if __name__ == "__main__":
    main()
