#!/usr/bin/env python
"""CUDA-event breakdown of one training step of the bench workload: forward graph, backward graph, optimizer, and the
same with the side lane disabled (how much of the weight-gradient work hides behind the dgrad chain).
usage (GPU box): python tools/step_breakdown.py [--batch 4]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vit_ae_plus_plus_b200.model import model_factory  # noqa: E402
from vit_ae_plus_plus_b200.utils import misc  # noqa: E402


def timed(fn, reps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    a = ap.parse_args()
    dev = torch.device("cuda")
    args = argparse.Namespace(model="mae_vit_base_patch16", volume_size=128, in_channels=4, patch_size=16,
                              perceptual_weight=0, use_imagenet=False)
    model = model_factory.get_models("autoenc", args).to(dev)
    model.pred_dtype = torch.bfloat16
    opt = torch.optim.AdamW(misc.add_weight_decay(model, 0.05), lr=1e-4, betas=(0.9, 0.95))
    scaler = misc.NativeScalerWithGradNormCount()
    x = torch.randn(a.batch, 4, 128, 128, 128, device=dev)
    eng = model.engine()

    def step():
        losses, _, _ = model(x, mask_ratio=0.75)
        scaler(losses[0], opt, parameters=model.parameters(), update_grad=True)
        opt.zero_grad()

    for side, overlap in ((True, True), (True, False), (False, False)):
        eng.use_side_lane = side
        scaler.overlap_optimizer = overlap
        for pl in eng.plans.values():
            pl.graphs.clear()
        for _ in range(4):
            step()
        noise = torch.rand(a.batch, eng.L, device=dev)
        keep = int(eng.L * 0.25)
        one = torch.ones(1, device=dev)
        t_step = timed(step)
        t_fwd = timed(lambda: eng.forward(x, noise, keep, pred_f32=False))
        pl = eng.forward(x, noise, keep, pred_f32=False)
        t_bwd = timed(lambda: eng.backward(pl, one, accumulate=False))
        fo = eng.fused_optimizer()
        t_opt = timed(lambda: fo.step(opt, scaler._scaler))
        eng.wait_params()
        print(f"side_lane={side} overlap_opt={overlap}: step {t_step:.3f} ms | forward {t_fwd:.3f} | backward {t_bwd:.3f} | optimizer {t_opt:.3f} "
              f"| sum {t_fwd + t_bwd + t_opt:.3f}", flush=True)


if __name__ == "__main__":
    main()
