#!/usr/bin/env python
"""Multi-GPU check of the data-parallel path (run under torchrun on N >= 2 GPUs of one box):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
1. the flat parameter buffer is identical on all ranks after construction (broadcast from rank 0, ranks seed differently);
2. backward(sync_grads=True) -- staged graphs with the per-stage NCCL all-reduce overlapped -- leaves in every rank's
   .grad the mean over ranks of the local gradients (compared with a no_sync backward + one explicit all-reduce);
3. after a few optimizer steps the replicas are still bit-identical."""
import argparse
import os
import sys
from functools import partial

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle import mae_oracle as O
    from vit_ae_plus_plus_b200.model.vit_autoenc import MaskedAutoencoderViT
    from vit_ae_plus_plus_b200.utils import misc
    cfg = O.CONFIGS["small"]
    torch.manual_seed(42 + rank)                       # different init per rank, as in the k-fold scripts
    m = MaskedAutoencoderViT(**cfg, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                             args=argparse.Namespace(perceptual_weight=0, use_imagenet=False)).to(dev)
    V, C = cfg["volume_size"], cfg["in_chans"]
    L = (V // cfg["patch_size"]) ** 3
    g = torch.Generator().manual_seed(100 + rank)      # different data per rank
    xs = [torch.randn(2, C, V, V, V, generator=g).to(dev) for _ in range(2)]
    noises = [torch.rand(2, L, generator=g) for _ in range(8)]
    eng = m.engine()

    def spread(t):
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return (hi - lo).abs().max().item()
    assert spread(eng.flat.p32) == 0.0, "parameters differ across ranks after the broadcast"

    # 2. staged + overlapped gradient exchange == local backward + explicit mean
    for rep in range(3):                               # eager, capture, replay
        with m.no_sync():
            losses, _, _ = m(xs[0], noise=noises[0])
            losses[0].backward()
        ref = eng.flat.g32.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.AVG)
        for p in m.parameters():
            p.grad = None
        losses, _, _ = m(xs[0], noise=noises[0])
        losses[0].backward()
        got = eng.flat.g32.clone()
        for p in m.parameters():
            p.grad = None
        err = (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-30)
        assert err < 1e-6, f"rep {rep}: staged all-reduce differs from the explicit mean by {err}"
        assert spread(got) == 0.0, "gradients differ across ranks after the exchange"

    # 3. replicas stay identical through optimizer steps
    opt = torch.optim.AdamW(misc.add_weight_decay(m, 0.05), lr=1e-3, betas=(0.9, 0.95))
    scaler = misc.NativeScalerWithGradNormCount()
    for i in range(6):
        losses, _, _ = m(xs[i % 2], noise=noises[i])
        scaler(losses[0], opt, parameters=m.parameters(), update_grad=True)
        opt.zero_grad()
    assert spread(eng.flat.p32) == 0.0, "replicas diverged"
    stages = len(eng._backward_stages(next(iter(eng.plans.values())), None, False, split=True))
    if rank == 0:
        print(f"dp_check ok: world {world}, {stages} backward stages, final loss {losses[0].item():.5f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
