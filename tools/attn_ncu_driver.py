#!/usr/bin/env python
"""Launches each attention kernel a few times on the decoder / encoder shapes of BASELINE configs[1] (for ncu)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vit_ae_plus_plus_b200 import ops  # noqa: E402

for B, N, H, hd in [(4, 513, 16, 32), (4, 129, 12, 64)]:
    D = H * hd
    qkv = torch.randn(B, N, 3 * D, device="cuda").bfloat16()
    dout = torch.randn(B, N, D, device="cuda").bfloat16()
    out = torch.empty(B, N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device="cuda")
    delta = torch.empty(B, H, N, device="cuda")
    dqkv = torch.empty(B, N, 3 * D, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.attention_fwd(qkv, out, lse, B, N, H, hd, hd ** -0.5)
        ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, B, N, H, hd, hd ** -0.5)
    torch.cuda.synchronize()
