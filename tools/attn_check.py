#!/usr/bin/env python
"""Diagnostic for the attention kernels (run on the GPU box): per shape, the error of out / lse / dq / dk / dv against the
fp32 torch reference of model/vit.py:112-121 and the device time of forward and backward (CUDA events, L2 flushed between
launches).  VITAE_ATTN_LEGACY=1 selects the round-1 mma.sync kernels for A/B.  Prints one line per shape; never asserts."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vit_ae_plus_plus_b200 import ops  # noqa: E402

DEV = "cuda"
SHAPES = [(4, 513, 16, 32), (4, 129, 12, 64), (2, 17, 4, 32), (2, 65, 4, 16), (1, 64, 2, 64), (3, 55, 16, 64),
          (1, 217, 16, 32), (1, 1, 1, 32), (1, 385, 12, 64), (1, 257, 12, 64), (2, 128, 3, 64), (2, 130, 2, 32),
          (16, 217, 16, 32), (16, 55, 16, 64)]


def rel(a, b):
    return (a.double() - b.double()).abs().max().item() / (b.double().abs().max().item() + 1e-30)


def main():
    tag = "legacy" if os.environ.get("VITAE_ATTN_LEGACY") == "1" else "tcgen05"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for B, N, H, hd in SHAPES:
        try:
            g = torch.Generator(device=DEV).manual_seed(N * 7 + hd)
            D = H * hd
            qkv = torch.randn(B, N, 3 * D, generator=g, device=DEV).bfloat16()
            dout = torch.randn(B, N, D, generator=g, device=DEV).bfloat16()
            out = torch.full((B, N, D), float("nan"), device=DEV, dtype=torch.bfloat16)
            lse = torch.full((B, H, N), float("nan"), device=DEV)
            delta = torch.empty(B, H, N, device=DEV)
            dqkv = torch.full((B, N, 3 * D), float("nan"), device=DEV, dtype=torch.bfloat16)
            scale = hd ** -0.5
            ops.attention_fwd(qkv, out, lse, B, N, H, hd, scale)
            ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, B, N, H, hd, scale)
            torch.cuda.synchronize()
            qr = qkv.float().requires_grad_(True)
            q, k, v = qr.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
            s = (q @ k.transpose(-2, -1)) * scale
            ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, N, D)
            lse_ref = torch.logsumexp(s, dim=-1)
            ref.backward(dout.float())
            gref = qr.grad.reshape(B, N, 3, D)
            got = dqkv.float().reshape(B, N, 3, D)
            errs = [rel(out.float(), ref.detach()), (lse - lse_ref.detach()).abs().max().item()] + \
                   [rel(got[:, :, i], gref[:, :, i]) if N > 1 else (got[:, :, i] - gref[:, :, i]).abs().max().item() for i in range(3)]
            tf, tb = [], []
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            for _ in range(6):
                flush.zero_()
                e0.record()
                ops.attention_fwd(qkv, out, lse, B, N, H, hd, scale)
                e1.record()
                ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, B, N, H, hd, scale)
                e2.record()
                torch.cuda.synchronize()
                tf.append(e0.elapsed_time(e1) * 1e3)
                tb.append(e1.elapsed_time(e2) * 1e3)
            nan = int(torch.isnan(out.float()).sum() + torch.isnan(dqkv.float()).sum())
            print(f"{tag} B={B} N={N} H={H} hd={hd}: out {errs[0]:.2e} lse {errs[1]:.2e} dq {errs[2]:.2e} dk {errs[3]:.2e} "
                  f"dv {errs[4]:.2e} nan {nan} | fwd {min(tf[1:]):.1f} us bwd {min(tb[1:]):.1f} us (cold L2)", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{tag} B={B} N={N} H={H} hd={hd}: FAILED {type(e).__name__}: {str(e)[:300]}", flush=True)
            break


if __name__ == "__main__":
    main()
