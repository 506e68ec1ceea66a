#!/usr/bin/env python
"""One tiny training step (forward + backward + fused AdamW), eager launches with the three backward lanes on, for
compute-sanitizer (memcheck / racecheck / synccheck).  CUDA graphs are off: the sanitizer instruments plain launches."""
import argparse
import os
import sys
from functools import partial

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vit_ae_plus_plus_b200.model.vit_autoenc import ContrastiveMAEViT  # noqa: E402
from vit_ae_plus_plus_b200.utils import misc  # noqa: E402
from vit_ae_plus_plus_b200.utils.train_one_epoch import compute_contrastive_loss  # noqa: E402

torch.manual_seed(0)
m = ContrastiveMAEViT(volume_size=32, patch_size=8, in_chans=2, embed_dim=128, depth=2, num_heads=4, decoder_embed_dim=64,
                      decoder_depth=2, decoder_num_heads=4, mlp_ratio=4, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                      args=argparse.Namespace(perceptual_weight=0, use_imagenet=False)).cuda().train()
m.use_cuda_graph = False
opt = torch.optim.AdamW(misc.add_weight_decay(m, 0.05), lr=1e-3, betas=(0.9, 0.95))
scaler = misc.NativeScalerWithGradNormCount()
crit = torch.nn.CosineSimilarity(dim=1)
args = argparse.Namespace(contr_weight=0.1)
raw = (torch.rand(2, 2, 32, 32, 32) * 4000).to(torch.float16).cuda()
for step in range(2):
    x = misc.normalize_volumes(raw, "z_score_channel")
    losses, pred, mask, p1, p2, z1, z2 = m(x, x.flip(2), mask_ratio=0.75, edge_map_weight=0.01)
    loss = losses[0] + compute_contrastive_loss(args, crit, p1, p2, z1, z2)
    scaler(loss, opt, parameters=m.parameters(), update_grad=True)
    opt.zero_grad()
torch.cuda.synchronize()
print("SANITIZE_STEP_OK", float(loss))
