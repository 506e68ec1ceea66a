"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's 3D ViT masked-autoencoder path.

This file is the checker for the CUDA product path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
(``vit_ae_plus_plus_b200/``) never does and fails loudly when its CUDA library is missing.

It is a *functional* restatement (plain tensors in a dict keyed by the reference's state_dict names, no
``nn.Module``) of what the reference computes, in torch CPU fp32 (or fp64 when the params are fp64).  The
arithmetic itself (conv/linear/layernorm/softmax/gelu) lives in torch -- a third-party dependency of the reference
pinned at ``torch==1.10.2+cu113`` (``requirements.txt:15``); here torch 2.11 runs it, same math.

Parity pin: the reference has NO tests or golden vectors of its own.  The oracle is therefore
pinned against outputs of the unmodified reference executed in the build container: ``oracle/make_golden.py``
-> ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` (plus a live comparison whenever ``/root/reference``
exists).

Every function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

# ----------------------------------------------------------------------------------------------------------------
# Configurations (BASELINE.json "configs"); decoder dims of "tiny" are our choice
# ----------------------------------------------------------------------------------------------------------------
CONFIGS = {
    # configs[0]: ViT-AE tiny (depth=2, dim=128, patch=8) on 32^3 x 1
    "tiny": dict(volume_size=32, patch_size=8, in_chans=1, embed_dim=128, depth=2, num_heads=4,
                 decoder_embed_dim=64, decoder_depth=2, decoder_num_heads=4, mlp_ratio=4),
    # a second small case with 2 channels / head_dim 32 in both stacks, ragged token counts
    "small": dict(volume_size=48, patch_size=16, in_chans=2, embed_dim=192, depth=2, num_heads=6,
                  decoder_embed_dim=128, decoder_depth=1, decoder_num_heads=4, mlp_ratio=4),
    # configs[1], [2], [4]: mae_vit_base_patch16 (model/vit_autoenc.py:296-301) on 128^3 x 4
    "vit_base_128": dict(volume_size=128, patch_size=16, in_chans=4, embed_dim=768, depth=12, num_heads=12,
                         decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4),
    # configs[3]: mae_vit_large_patch16 (model/vit_autoenc.py:288-293) on 96^3 x 4
    "vit_large_96": dict(volume_size=96, patch_size=16, in_chans=4, embed_dim=1024, depth=24, num_heads=16,
                         decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4),
}
LN_EPS = 1e-6  # model/vit_autoenc.py:292,300 (partial(nn.LayerNorm, eps=1e-6))


def geometry(cfg):
    g = cfg["volume_size"] // cfg["patch_size"]
    L = g ** 3
    P = cfg["patch_size"] ** 3 * cfg["in_chans"]
    return g, L, P


# ----------------------------------------------------------------------------------------------------------------
# Positional embedding -- model/model_utils/vit_helpers.py:13-70
# ----------------------------------------------------------------------------------------------------------------
def _sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    # vit_helpers.py:48-70: omega_k = 10000^(-k/(dim/2)); [sin | cos]
    assert dim % 2 == 0
    omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=float) / (dim / 2.0))
    ang = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def sincos_pos_embed_3d(embed_dim: int, grid_size: int, cls_token: bool = True) -> np.ndarray:
    # vit_helpers.py:19-29: np.meshgrid with default 'xy' indexing -> grid[0] varies along axis 1, grid[1] along
    # axis 0, grid[2] along axis 2 of the (g,g,g) raster.  vit_helpers.py:36-39: res = ceil_even(D//3).
    ar = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(ar, ar, ar), axis=0).reshape(3, 1, grid_size, grid_size, grid_size)
    res = embed_dim // 3
    if res % 2:
        res += 1
    last = embed_dim - 2 * res
    emb = np.concatenate([_sincos_1d(res, grid[0]), _sincos_1d(res, grid[1]), _sincos_1d(last, grid[2])], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)  # vit_helpers.py:27-28
    return emb


# ----------------------------------------------------------------------------------------------------------------
# Parameters -- names/shapes of model/vit_autoenc.py:18-63, init of :65-98
# ----------------------------------------------------------------------------------------------------------------
def param_shapes(cfg) -> Dict[str, tuple]:
    g, L, P = geometry(cfg)
    D, Dd, C, p = cfg["embed_dim"], cfg["decoder_embed_dim"], cfg["in_chans"], cfg["patch_size"]
    hid, hidd = int(D * cfg["mlp_ratio"]), int(Dd * cfg["mlp_ratio"])
    s = {"cls_token": (1, 1, D), "pos_embed": (1, L + 1, D),
         "patch_embed.proj.weight": (D, C, p, p, p), "patch_embed.proj.bias": (D,)}

    def block(prefix, d, h):
        s[f"{prefix}.norm1.weight"] = (d,); s[f"{prefix}.norm1.bias"] = (d,)
        s[f"{prefix}.attn.qkv.weight"] = (3 * d, d); s[f"{prefix}.attn.qkv.bias"] = (3 * d,)
        s[f"{prefix}.attn.proj.weight"] = (d, d); s[f"{prefix}.attn.proj.bias"] = (d,)
        s[f"{prefix}.norm2.weight"] = (d,); s[f"{prefix}.norm2.bias"] = (d,)
        s[f"{prefix}.mlp.fc1.weight"] = (h, d); s[f"{prefix}.mlp.fc1.bias"] = (h,)
        s[f"{prefix}.mlp.fc2.weight"] = (d, h); s[f"{prefix}.mlp.fc2.bias"] = (d,)

    for i in range(cfg["depth"]):
        block(f"blocks.{i}", D, hid)
    s["norm.weight"] = (D,); s["norm.bias"] = (D,)
    s["decoder_embed.weight"] = (Dd, D); s["decoder_embed.bias"] = (Dd,)
    s["mask_token"] = (1, 1, Dd); s["decoder_pos_embed"] = (1, L + 1, Dd)
    for i in range(cfg["decoder_depth"]):
        block(f"decoder_blocks.{i}", Dd, hidd)
    s["decoder_norm.weight"] = (Dd,); s["decoder_norm.bias"] = (Dd,)
    s["decoder_pred.weight"] = (P, Dd); s["decoder_pred.bias"] = (P,)
    return s


FROZEN = ("pos_embed", "decoder_pos_embed")  # requires_grad=False, vit_autoenc.py:30-31,45-46


def init_params(cfg, seed: int = 0, dtype=torch.float32, perturb: float = 0.02) -> Params:
    """Deterministic parameters with the reference's init *distributions* (vit_autoenc.py:65-98): xavier-uniform
    for every Linear and for patch_embed.proj.weight viewed (D,-1), N(0,.02) cls/mask tokens, sin-cos frozen
    pos-embeds.  Biases / LayerNorm affine are perturbed by ``perturb``*N(0,1) around the reference's (0, 1) so
    that parity tests exercise them (the reference init leaves them exactly 0 / 1)."""
    gen = torch.Generator().manual_seed(seed)
    g, L, P = geometry(cfg)
    out: Params = {}
    for name, shape in param_shapes(cfg).items():
        if name == "pos_embed":
            t = torch.from_numpy(sincos_pos_embed_3d(cfg["embed_dim"], g)).float().unsqueeze(0)
        elif name == "decoder_pos_embed":
            t = torch.from_numpy(sincos_pos_embed_3d(cfg["decoder_embed_dim"], g)).float().unsqueeze(0)
        elif name in ("cls_token", "mask_token"):
            t = torch.randn(shape, generator=gen) * 0.02
        elif name.endswith("norm1.weight") or name.endswith("norm2.weight") or name in ("norm.weight", "decoder_norm.weight"):
            t = 1.0 + perturb * torch.randn(shape, generator=gen)
        elif len(shape) == 1:
            t = perturb * torch.randn(shape, generator=gen)
        else:
            fan_out, fan_in = shape[0], int(np.prod(shape[1:]))
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=gen) * 2 - 1) * bound
        out[name] = t.to(dtype)
    return out


# ----------------------------------------------------------------------------------------------------------------
# Forward pieces
# ----------------------------------------------------------------------------------------------------------------
def patch_embed(x, w, b):
    # model/vit.py:68-76: Conv3d(k=s=p) -> flatten(2).transpose(1,2); token order raster (d,h,w)
    p = w.shape[-1]
    return F.conv3d(x, w, b, stride=p).flatten(2).transpose(1, 2)


def random_masking(x, mask_ratio: float, noise: torch.Tensor):
    # model/vit_autoenc.py:130-155 with the noise drawn by the caller (reference: torch.rand(N, L) at :139)
    N, L, D = x.shape
    len_keep = int(L * (1 - mask_ratio))                                   # :137
    ids_shuffle = torch.argsort(noise, dim=1)                              # :142
    ids_restore = torch.argsort(ids_shuffle, dim=1)                        # :143
    ids_keep = ids_shuffle[:, :len_keep]                                   # :146
    x_masked = torch.gather(x, 1, ids_keep.unsqueeze(-1).expand(-1, -1, D))  # :147
    mask = torch.ones(N, L, dtype=x.dtype, device=x.device)
    mask[:, :len_keep] = 0                                                 # :150-151
    mask = torch.gather(mask, 1, ids_restore)                              # :153
    return x_masked, mask, ids_restore


def attention(x, P: Params, prefix: str, num_heads: int):
    # model/vit.py:112-124
    B, N, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, P[f"{prefix}.qkv.weight"], P[f"{prefix}.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    attn = attn.softmax(dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, P[f"{prefix}.proj.weight"], P[f"{prefix}.proj.bias"])


def block(x, P: Params, prefix: str, num_heads: int):
    # model/vit.py:139-144 (+ Mlp3D :90-96, nn.GELU default = erf form)
    C = x.shape[-1]
    h = F.layer_norm(x, (C,), P[f"{prefix}.norm1.weight"], P[f"{prefix}.norm1.bias"], LN_EPS)
    x = x + attention(h, P, f"{prefix}.attn", num_heads)
    h = F.layer_norm(x, (C,), P[f"{prefix}.norm2.weight"], P[f"{prefix}.norm2.bias"], LN_EPS)
    h = F.linear(h, P[f"{prefix}.mlp.fc1.weight"], P[f"{prefix}.mlp.fc1.bias"])
    h = F.gelu(h)
    h = F.linear(h, P[f"{prefix}.mlp.fc2.weight"], P[f"{prefix}.mlp.fc2.bias"])
    return x + h


def forward_encoder(x, P: Params, cfg, mask_ratio: float, noise: torch.Tensor):
    # model/vit_autoenc.py:157-177
    t = patch_embed(x, P["patch_embed.proj.weight"], P["patch_embed.proj.bias"])
    t = t + P["pos_embed"][:, 1:, :]
    t, mask, ids_restore = random_masking(t, mask_ratio, noise)
    cls = (P["cls_token"] + P["pos_embed"][:, :1, :]).expand(t.shape[0], -1, -1)
    t = torch.cat([cls, t], dim=1)
    for i in range(cfg["depth"]):
        t = block(t, P, f"blocks.{i}", cfg["num_heads"])
    D = t.shape[-1]
    t = F.layer_norm(t, (D,), P["norm.weight"], P["norm.bias"], LN_EPS)
    return t, mask, ids_restore


def forward_decoder(latent, P: Params, cfg, ids_restore):
    # model/vit_autoenc.py:179-203
    x = F.linear(latent, P["decoder_embed.weight"], P["decoder_embed.bias"])
    B, Ne, Dd = x.shape
    L = ids_restore.shape[1]
    mask_tokens = P["mask_token"].expand(B, L + 1 - Ne, -1)
    x_ = torch.cat([x[:, 1:, :], mask_tokens], dim=1)
    x_ = torch.gather(x_, 1, ids_restore.unsqueeze(-1).expand(-1, -1, Dd))
    x = torch.cat([x[:, :1, :], x_], dim=1)
    x = x + P["decoder_pos_embed"]
    for i in range(cfg["decoder_depth"]):
        x = block(x, P, f"decoder_blocks.{i}", cfg["decoder_num_heads"])
    x = F.layer_norm(x, (Dd,), P["decoder_norm.weight"], P["decoder_norm.bias"], LN_EPS)
    x = F.linear(x, P["decoder_pred.weight"], P["decoder_pred.bias"])
    return x[:, 1:, :]


def patchify(vol, p: int):
    # model/vit_autoenc.py:100-113 -- within-patch order (pz, py, px, c), channel fastest
    N, C, V = vol.shape[0], vol.shape[1], vol.shape[2]
    assert vol.shape[2] == vol.shape[3] == vol.shape[4] and V % p == 0
    g = V // p
    x = vol.reshape(N, C, g, p, g, p, g, p)
    x = x.permute(0, 2, 4, 6, 3, 5, 7, 1)       # n l h w r p q c
    return x.reshape(N, g * g * g, p * p * p * C)


def unpatchify(x, p: int):
    # model/vit_autoenc.py:115-128
    N, L = x.shape[0], x.shape[1]
    g = round(L ** (1 / 3))
    assert g * g * g == L
    x = x.reshape(N, g, g, g, p, p, p, -1)
    x = x.permute(0, 7, 1, 4, 2, 5, 3, 6)       # n c l r h p w q
    return x.reshape(N, -1, g * p, g * p, g * p)


def masked_mse(pred, target, mask):
    # model/vit_autoenc.py:226-227: mean over P, weight by mask, sum / mask.sum() over the whole batch
    per_patch = ((pred - target) ** 2).mean(dim=-1)
    return (per_patch * mask).sum() / mask.sum()


# -- auxiliary edge-map term -------------------------------------------------------------
def sobel_kernels(dtype=torch.float32):
    # model/model_utils/sobel_filter.py:10-35: smoothing [1,2,1] x [1,2,1] x derivative [1,0,-1] / [-1,0,1]
    s = torch.tensor([1.0, 2.0, 1.0], dtype=dtype)
    d = torch.tensor([1.0, 0.0, -1.0], dtype=dtype)
    k0 = torch.einsum("i,j,k->ijk", s, s, d)       # weight[0,0]: derivative along the last axis, sign (+,0,-)
    k1 = torch.einsum("i,j,k->ijk", s, -d, s)      # weight[1,0]: derivative along the middle axis (-,0,+)
    k2 = torch.einsum("i,j,k->ijk", -d, s, s)      # weight[2,0]: derivative along the first axis (-,0,+)
    return torch.stack([k0, k1, k2]).unsqueeze(1)  # (3,1,3,3,3)


def sobel_edge_map(vol):
    # model/model_utils/sobel_filter.py:37-45: per channel sqrt(gx^2+gy^2+gz^2), summed over channels
    w = sobel_kernels(vol.dtype)
    out = 0
    for c in range(vol.shape[1]):
        g = F.conv3d(vol[:, c:c + 1], w, None, stride=1, padding=1)
        out = out + torch.sqrt((g ** 2).sum(dim=1))
    return out


def gaussian_taps(sigma=2.0, dtype=torch.float32):
    # model/model_utils/gaussian_filter.py:5-13 incl. the linspace(-ks//2, ks//2+1, ks) quirk (taps 1.2 apart)
    ks = int(sigma * 5)
    if ks % 2 == 0:
        ks += 1
    ts = torch.linspace(-ks // 2, ks // 2 + 1, ks, dtype=dtype)
    gk = torch.exp(-(ts / sigma) ** 2 / 2)
    return gk / gk.sum()


def gaussian_blur_3d(vol, sigma=2.0):
    # model/model_utils/gaussian_filter.py:16-26: dense ks^3 kernel, renormalised, per channel, zero padding
    k = gaussian_taps(sigma, vol.dtype)
    k3 = torch.einsum("i,j,k->ijk", k, k, k)
    k3 = (k3 / k3.sum()).reshape(1, 1, *k3.shape)
    outs = [F.conv3d(vol[:, c:c + 1], k3, stride=1, padding=len(k) // 2) for c in range(vol.shape[1])]
    return torch.cat(outs, dim=1)


def edge_map_mse(pred, target, p: int):
    # model/vit_autoenc.py:221-224
    pv, tv = unpatchify(pred, p), unpatchify(target, p)
    return F.mse_loss(sobel_edge_map(pv), sobel_edge_map(gaussian_blur_3d(tv, 2.0)), reduction="mean")


def forward(x, P: Params, cfg, mask_ratio: float, noise: torch.Tensor, edge_map_weight: float = 0.0,
            with_edge: bool = True):
    """model/vit_autoenc.py:234-238 + forward_loss :205-232 with perceptual_weight = 0 (the shipped default,
    config.ini:34; the VGG term is then exactly 0 and never differentiable, perceptual_loss.py:68-69).
    Returns ([loss, raw_edge, recon, percep], pred, mask, ids_restore)."""
    latent, mask, ids_restore = forward_encoder(x, P, cfg, mask_ratio, noise)
    pred = forward_decoder(latent, P, cfg, ids_restore)
    target = patchify(x, cfg["patch_size"])
    recon = masked_mse(pred, target, mask)
    if with_edge:
        raw_edge = edge_map_mse(pred, target, cfg["patch_size"])
    else:
        raw_edge = torch.zeros((), dtype=pred.dtype, device=pred.device)
    percep = torch.zeros((), dtype=pred.dtype, device=pred.device)
    loss = edge_map_weight * raw_edge + recon + percep
    return [loss, raw_edge, recon, percep], pred, mask, ids_restore


def forward_backward(x, P: Params, cfg, mask_ratio: float, noise: torch.Tensor, edge_map_weight: float = 0.0,
                     with_edge: bool = False):
    """Runs forward + autograd backward of loss[0]; returns (losses, pred, mask, grads dict)."""
    leaves = {k: v.detach().clone().requires_grad_(k not in FROZEN) for k, v in P.items()}
    losses, pred, mask, _ = forward(x, leaves, cfg, mask_ratio, noise, edge_map_weight, with_edge)
    losses[0].backward()
    grads = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    return [l.detach() for l in losses], pred.detach(), mask, grads


# ----------------------------------------------------------------------------------------------------------------
# Contrastive wrapper -- model/vit_autoenc.py:241-285 (ContrastiveMAEViT, use_proj=False) and the loss of
# utils/train_one_epoch.py:113-114
# ----------------------------------------------------------------------------------------------------------------
BN_EPS = 1e-5  # nn.BatchNorm1d default (vit_autoenc.py:265)


def predictor_param_shapes(cfg) -> Dict[str, tuple]:
    D = cfg["embed_dim"]
    return {"predictor.0.weight": (D, D), "predictor.1.weight": (D,), "predictor.1.bias": (D,),
            "predictor.3.weight": (D, D), "predictor.3.bias": (D,)}


def init_predictor_params(cfg, seed: int = 0, dtype=torch.float32, perturb: float = 0.02) -> Params:
    """The predictor is built after super().__init__() and keeps torch's default Linear / BatchNorm init
    (vit_autoenc.py:263-268); here: the same kaiming-uniform(a=sqrt(5)) range for the weights, perturbed
    affine / bias so that tests exercise them."""
    gen = torch.Generator().manual_seed(seed + 7919)
    out: Params = {}
    for name, shape in predictor_param_shapes(cfg).items():
        if len(shape) == 2:
            bound = 1.0 / math.sqrt(shape[1])
            t = (torch.rand(shape, generator=gen) * 2 - 1) * bound
        elif name == "predictor.1.weight":
            t = 1.0 + perturb * torch.randn(shape, generator=gen)
        else:
            t = perturb * torch.randn(shape, generator=gen)
        out[name] = t.to(dtype)
    return out


def predictor(z, P: Params):
    """Linear(no bias) -> BatchNorm1d in training mode (batch statistics over the B*Ne token rows, biased variance)
    -> ReLU -> Linear   (vit_autoenc.py:263-268)."""
    h = F.linear(z, P["predictor.0.weight"])
    mu, var = h.mean(0), h.var(0, unbiased=False)
    h = (h - mu) / torch.sqrt(var + BN_EPS) * P["predictor.1.weight"] + P["predictor.1.bias"]
    h = F.relu(h)
    return F.linear(h, P["predictor.3.weight"], P["predictor.3.bias"])


def forward_contrastive(x1, x2, P: Params, cfg, mask_ratio: float, noise1: torch.Tensor, noise2: torch.Tensor,
                        edge_map_weight: float = 0.0, with_edge: bool = False):
    """vit_autoenc.py:270-285 -> (loss_list, pred, mask, p1, p2, z1.detach(), z2.detach()); the second view is masked
    with its own noise draw (:277)."""
    latent1, mask, ids_restore = forward_encoder(x1, P, cfg, mask_ratio, noise1)
    pred = forward_decoder(latent1, P, cfg, ids_restore)
    target = patchify(x1, cfg["patch_size"])
    recon = masked_mse(pred, target, mask)
    raw_edge = edge_map_mse(pred, target, cfg["patch_size"]) if with_edge else torch.zeros((), dtype=pred.dtype, device=pred.device)
    percep = torch.zeros((), dtype=pred.dtype, device=pred.device)
    latent2, _, _ = forward_encoder(x2, P, cfg, mask_ratio, noise2)
    z1 = latent1.reshape(-1, latent1.shape[2])
    z2 = latent2.reshape(-1, latent2.shape[2])
    return ([edge_map_weight * raw_edge + recon + percep, raw_edge, recon, percep], pred, mask, predictor(z1, P),
            predictor(z2, P), z1.detach(), z2.detach())


def contrastive_loss(p1, p2, z1, z2, contr_weight: float):
    # utils/train_one_epoch.py:113-114 with criterion = nn.CosineSimilarity(dim=1) (:32)
    cos = lambda a, b: F.cosine_similarity(a, b, dim=1)
    return contr_weight * (-(cos(p1, z2).mean() + cos(p2, z1).mean()) * 0.5)


# ----------------------------------------------------------------------------------------------------------------
# Encoder-only feature extraction -- model/vit.py:265-284 (VisionTransformer3D.forward_features),
# the model the k-fold scripts load the MAE checkpoint into (k_fold_cross_valid_combined_brats.py:219-245)
# ----------------------------------------------------------------------------------------------------------------
def vit_param_names(cfg, global_pool: bool):
    """state_dict keys of VisionTransformer3D(embed_dim/depth/heads of cfg): the MAE encoder's keys + the head (+ fc_norm
    instead of norm with global_pool, model/vit.py:219-222)."""
    names = ["cls_token", "pos_embed", "patch_embed.proj.weight", "patch_embed.proj.bias"]
    for i in range(cfg["depth"]):
        for n in ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
                  "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias"):
            names.append(f"blocks.{i}.{n}")
    names += ["fc_norm.weight", "fc_norm.bias"] if global_pool else ["norm.weight", "norm.bias"]
    return names + ["head.weight", "head.bias"]


def init_vit_params(cfg, num_classes: int, global_pool: bool, seed: int = 0, perturb: float = 0.02) -> Params:
    """Encoder parameters as init_params() produces them (what a pre-trained MAE checkpoint carries) + head / fc_norm."""
    P = init_params(cfg, seed)
    gen = torch.Generator().manual_seed(seed + 4242)
    D = cfg["embed_dim"]
    out = {k: v for k, v in P.items() if k in set(vit_param_names(cfg, global_pool))}
    if global_pool:
        out["fc_norm.weight"] = 1.0 + perturb * torch.randn(D, generator=gen)
        out["fc_norm.bias"] = perturb * torch.randn(D, generator=gen)
    out["head.weight"] = 0.02 * torch.randn(num_classes, D, generator=gen)
    out["head.bias"] = perturb * torch.randn(num_classes, generator=gen)
    return out


def vit_forward_features(x, P: Params, cfg, global_pool: bool):
    # model/vit.py:265-284: all patches (no masking), cls + pos, blocks, then mean-pool + fc_norm or norm + cls row
    t = patch_embed(x, P["patch_embed.proj.weight"], P["patch_embed.proj.bias"])
    t = torch.cat([P["cls_token"].expand(t.shape[0], -1, -1), t], dim=1) + P["pos_embed"]
    for i in range(cfg["depth"]):
        t = block(t, P, f"blocks.{i}", cfg["num_heads"])
    D = t.shape[-1]
    if global_pool:
        return F.layer_norm(t[:, 1:, :].mean(dim=1), (D,), P["fc_norm.weight"], P["fc_norm.bias"], LN_EPS)
    return F.layer_norm(t, (D,), P["norm.weight"], P["norm.bias"], LN_EPS)[:, 0]


def vit_forward(x, P: Params, cfg, global_pool: bool):
    # model/vit.py:286-297 (no distillation head)
    return F.linear(vit_forward_features(x, P, cfg, global_pool), P["head.weight"], P["head.bias"])


# ----------------------------------------------------------------------------------------------------------------
# Intensity normalisation of an incoming volume [C, V, V, V] -- Dataset._normalize_data
# ----------------------------------------------------------------------------------------------------------------
def normalize_volume(volume: torch.Tensor, mode: str) -> torch.Tensor:
    volume = volume.float()
    if mode == "z_score_channel":      # dataset/egd_dataset/egd.py:45-47 (torch.var: unbiased)
        return (volume - torch.mean(volume, dim=[1, 2, 3], keepdim=True)) / torch.sqrt(torch.var(volume, dim=[1, 2, 3], keepdim=True))
    if mode == "z_score_sample":       # dataset/brats_dataset/brats.py:27-29
        return (volume - volume.mean()) / torch.sqrt(volume.var())
    if mode == "min_max":              # egd.py:48-50, brats.py:30-32
        max_val, min_val = volume.max(), volume.min()
        return 2 * ((volume - min_val) / (max_val - min_val)) - 1
    raise ValueError(mode)


# ----------------------------------------------------------------------------------------------------------------
# Optimizer grouping + AdamW as built at the call site (k_fold_cross_valid_combined_brats.py:168-169)
# ----------------------------------------------------------------------------------------------------------------
def weight_decay_groups(named_params, weight_decay: float):
    """timm 0.5.4 optim_factory.add_weight_decay semantics: frozen skipped; ndim==1 or name ends
    with '.bias' -> no decay; everything else (incl. the 3-D cls/mask tokens) decayed."""
    decay, no_decay = [], []
    for name, p in named_params:
        if not p.requires_grad:
            continue
        (no_decay if (p.ndim == 1 or name.endswith(".bias")) else decay).append(p)
    return [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": weight_decay}]


def cosine_lr(epoch_float: float, lr: float, min_lr: float, warmup_epochs: float, epochs: float) -> float:
    # utils/lr_sched.py:9-21
    if epoch_float < warmup_epochs:
        return lr * epoch_float / warmup_epochs
    return min_lr + (lr - min_lr) * 0.5 * (1.0 + math.cos(math.pi * (epoch_float - warmup_epochs) / (epochs - warmup_epochs)))


def flops_per_volume(cfg, mask_ratio: float, kept_only_embed: bool = True):
    """Algorithmic GEMM FLOPs per volume: returns (F_fwd, F_step)."""
    g, L, P = geometry(cfg)
    D, Dd = cfg["embed_dim"], cfg["decoder_embed_dim"]
    keep = int(L * (1 - mask_ratio))
    Ne, Nd = keep + 1, L + 1
    embed = 2 * (keep if kept_only_embed else L) * P * D
    f_fwd = (embed + cfg["depth"] * (24 * Ne * D * D + 4 * Ne * Ne * D) + 2 * Ne * D * Dd
             + cfg["decoder_depth"] * (24 * Nd * Dd * Dd + 4 * Nd * Nd * Dd) + 2 * Nd * Dd * P)
    return f_fwd, 3 * f_fwd - embed
