#!/usr/bin/env python
"""Host-side cost of one training step (python + launches): wall time per step with the device idle-free (queue
never drains) vs the device time, and a cProfile of the step function.  usage (GPU box): python tools/host_profile.py"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vit_ae_plus_plus_b200.model import model_factory  # noqa: E402
from vit_ae_plus_plus_b200.utils import misc  # noqa: E402


def main():
    dev = torch.device("cuda")
    args = argparse.Namespace(model="mae_vit_base_patch16", volume_size=128, in_channels=4, patch_size=16,
                              perceptual_weight=0, use_imagenet=False)
    model = model_factory.get_models("autoenc", args).to(dev)
    opt = torch.optim.AdamW(misc.add_weight_decay(model, 0.05), lr=1e-4, betas=(0.9, 0.95))
    scaler = misc.NativeScalerWithGradNormCount()
    x = torch.randn(4, 4, 128, 128, 128, device=dev)

    def step():
        losses, _, _ = model(x, mask_ratio=0.75)
        scaler(losses[0], opt, parameters=model.parameters(), update_grad=True)
        opt.zero_grad()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    n = 50
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"host enqueue {1e3 * t_enq / n:.3f} ms/step, wall incl. device {1e3 * t_all / n:.3f} ms/step "
          f"(cpus: {os.cpu_count()})")
    # host-only cost: same loop while the device is kept far behind is not observable directly; profile instead
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
