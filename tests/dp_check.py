#!/usr/bin/env python
"""Multi-GPU check of the data-parallel path, one process per GPU over NCCL (tests/test_dp_gpu.py launches it under
torchrun when the box has >= 2 GPUs; bench.py runs the same check, dp.replica_check, inside every N > 1 run):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py
"""
import argparse
import json
import os
import sys
from functools import partial

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle import mae_oracle as O
    from vit_ae_plus_plus_b200 import dp
    from vit_ae_plus_plus_b200.model.vit_autoenc import MaskedAutoencoderViT
    cfg = O.CONFIGS["small"]
    torch.manual_seed(42 + rank)                       # different init per rank, as in the k-fold scripts
    m = MaskedAutoencoderViT(**cfg, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                             args=argparse.Namespace(perceptual_weight=0, use_imagenet=False)).to(dev)
    V, C = cfg["volume_size"], cfg["in_chans"]
    L = (V // cfg["patch_size"]) ** 3
    g = torch.Generator().manual_seed(100 + rank)      # different data per rank
    xs = [torch.randn(2, C, V, V, V, generator=g).to(dev) for _ in range(2)]
    noises = [torch.rand(2, L, generator=g) for _ in range(8)]
    res = dp.replica_check(m, xs, noises, opt_steps=6)
    assert res["ok"], res
    if rank == 0:
        print("DP_CHECK_OK " + json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
