#!/usr/bin/env python
"""Times the three edge-map loss entry points (target, forward, backward) separately at the headline shape
(B=4, C=4, V=128, p=16): CUDA graph of REPS launches each, CUDA events; prints us per call and the volume-passes
equivalent (time / time to stream one fp32 volume set at the measured HBM rate).
usage (GPU box): python tools/edge_bench.py > gpurun_out/edge_bench.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vit_ae_plus_plus_b200 import ops  # noqa: E402

REPS = 10


def timed(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * REPS)


def main():
    B, C, V, p = 4, 4, 128, 16
    dev = torch.device("cuda")
    L = (V // p) ** 3
    vol = torch.randn(B, C, V, V, V, device=dev)
    pred = torch.randn(B, L + 1, p ** 3 * C, device=dev).bfloat16()
    dpred = torch.zeros_like(pred)
    scratch = torch.empty(ops.edge_scratch_floats(B, C, V), device=dev)
    tgt = torch.empty(B, V, V, V, device=dev)
    resid = torch.empty(B, V, V, V, device=dev)
    out = torch.zeros(1, device=dev)
    up = torch.ones(1, device=dev)
    taps = ops.gaussian_taps(2.0)
    vol_bytes = vol.numel() * 4
    rows = [("edge_target   (3 blur passes + sobel)", lambda: ops.edge_target(vol, taps, scratch, tgt)),
            ("edge_loss_fwd (unpatchify + sobel + finalize)", lambda: ops.edge_loss_fwd(pred, tgt, scratch, resid, out, B, C, V, p)),
            ("edge_loss_bwd (transposed stencil)", lambda: ops.edge_loss_bwd(resid, scratch, up, dpred, B, C, V, p))]
    if "--eager" in sys.argv:          # one eager call each: for `ncu --set full` captures
        for _, fn in rows:
            fn()
        torch.cuda.synchronize()
        return
    total = 0.0
    for name, fn in rows:
        us = timed(fn)
        total += us
        print(f"{name:48s} {us:8.1f} us   = {us * 1e-6 * 6.5e12 / vol_bytes:5.1f} volume passes at 6.5 TB/s")
    print(f"{'total':48s} {total:8.1f} us   (volume = {vol_bytes / 1e6:.0f} MB fp32)")


if __name__ == "__main__":
    main()
