#!/usr/bin/env python
"""Timeline of the tcgen05 attention forward (debug build:
  VITAE_ATTN_TRACE=1 VITAE_BUILD_VARIANT=attntrace python -m vit_ae_plus_plus_b200.build
  VITAE_LIB=vit_ae_plus_plus_b200/libvitae_b200_attntrace.so python tools/attn_trace.py).
Per CTA: %globaltimer (ns) at 0 entry, 1 setup done, 2 Q landed (MMA thread), 3 pass-1 MMAs issued, 4 pass 1 done (softmax
warp), 5 last PV issued, 6 pass 2 done, 7 O ready, 8 CTA done; 9 SM id; accumulated barrier-wait cycles of one softmax thread
(10 s_full pass 1, 11 s_full pass 2, 12 p_free) and of the MMA thread (13 s_free, 14 p_full, 15 kv_full)."""
import ctypes
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vit_ae_plus_plus_b200 import _lib, ops  # noqa: E402

NAMES = ["entry", "setup", "q_landed", "s_ready|p1_issued", "p1_done", "pv_issued", "p2_done", "o_ready", "done"]


def main():
    dev = torch.device("cuda")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    trace_all = torch.zeros((1 << 18) + 3 * 512, dtype=torch.int64, device=dev)      # per-CTA slots, then CTA 0's event log
    trace = trace_all[:1 << 18].view(1 << 14, 16)
    lib.vitae_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
    assert lib.vitae_debug_set_attn_trace(trace_all.data_ptr()) == 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for B, N, H, hd in [(4, 513, 16, 32), (4, 512, 16, 32), (2, 512, 16, 32), (1, 512, 16, 32), (4, 129, 12, 64), (16, 217, 16, 32)]:
        D = H * hd
        qkv = torch.randn(B, N, 3 * D, device=dev).bfloat16()
        out = torch.empty(B, N, D, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, H, N, device=dev)
        for mode in ("warm", "cold"):
            for rep in range(3):
                if mode == "cold":
                    flush.zero_()
                trace.zero_()
                torch.cuda.synchronize()
                ops.attention_fwd(qkv, out, lse, B, N, H, hd, hd ** -0.5)
                torch.cuda.synchronize()
            t = trace.cpu()
            t = t[t[:, 0] > 0]
            t0 = int(t[:, 0].min())
            span = (int(t[:, 8].max()) - t0) / 1000
            dur = [(int(r[8]) - int(r[0])) / 1000 for r in t]
            per_sm = {}
            for r in t:
                per_sm.setdefault(int(r[9]), []).append(r)
            rel = ["%s=%.2f" % (n, statistics.median([(int(r[k]) - int(r[0])) / 1000 for r in t if r[k] > 0])) for k, n in enumerate(NAMES) if k > 0]
            waits = ["%s=%.0f" % (n, statistics.median([int(r[k]) for r in t])) for k, n in
                     [(10, "sm.s_full1"), (11, "sm.s_full2"), (12, "sm.p_free"), (13, "mma.s_free"), (14, "mma.p_full"), (15, "mma.kv_full")]]
            print(f"B={B} N={N} H={H} hd={hd} {mode}: ctas={t.shape[0]} sms={len(per_sm)} max_ctas_per_sm={max(len(v) for v in per_sm.values())} "
                  f"span={span:.2f} us  cta_dur median={statistics.median(dur):.2f} max={max(dur):.2f} | since entry (us, median): "
                  + " ".join(rel) + " | wait cycles (median): " + " ".join(waits), flush=True)


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def events(B=1, N=512, H=16, hd=32):
    """Event log of CTA 0 (VITAE_ATTN_FWD=v1): (id, cycle) pairs of the MMA thread, one softmax thread and the TMA thread."""
    dev = torch.device("cuda")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    trace = torch.zeros((1 << 18) + 3 * 512, dtype=torch.int64, device=dev)
    lib.vitae_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
    assert lib.vitae_debug_set_attn_trace(trace.data_ptr()) == 0
    D = H * hd
    qkv = torch.randn(B, N, 3 * D, device=dev).bfloat16()
    out = torch.empty(B, N, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, N, device=dev)
    for rep in range(3):
        trace.zero_()
        torch.cuda.synchronize()
        ops.attention_fwd(qkv, out, lse, B, N, H, hd, hd ** -0.5)
        torch.cuda.synchronize()
    ev = trace[1 << 18:].cpu().view(3, 256, 2)
    t0 = min(int(ev[w, 0, 1]) for w in range(3) if ev[w, 0, 1] > 0)
    for w, name in enumerate(["mma", "softmax", "tma"]):
        row = [(int(i), int(c) - t0) for i, c in ev[w].tolist() if c > 0]
        print(f"events {name} B={B} N={N} hd={hd} cfg={os.environ.get('VITAE_ATTN_FWD_CFG', '1')}: " + " ".join(f"{i}@{c}" for i, c in row), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "events":
    events()
    events(B=4, N=513)
