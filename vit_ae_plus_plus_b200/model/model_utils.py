"""Init-time host helpers of the MAE path: the frozen 3-D sin-cos positional table and the auxiliary edge-map loss.

* ``sincos_pos_embed_3d`` follows the reference's ``get_3d_sincos_pos_embed`` (model/model_utils/vit_helpers.py:13-70),
  including its quirks: ``np.meshgrid`` default 'xy' indexing (so the first channel third encodes the H index, the
  second the D index, the last the W index), channel split ``res = ceil_even(D // 3)`` with the remainder on the last
  axis, and an all-zero cls row.  fp64 numpy, cast to fp32 by the caller -- init only, never on the step.
* ``EdgeMapLoss`` is the interim implementation of SURVEY.md row f-1 (Sobel edge-map MSE against the Gaussian-blurred
  target, model/vit_autoenc.py:221-224): it is NOT part of the fused hot path and runs on torch convolution ops on
  the kernels' ``pred`` output until the fused stencil kernel exists.  It is only evaluated when
  ``edge_map_weight != 0`` (or when the module is asked to report the raw edge loss).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    if dim % 2:
        raise ValueError("sin-cos embedding needs an even channel count per axis")
    freq = np.power(10000.0, -np.arange(dim // 2, dtype=np.float64) / (dim / 2.0))
    ang = pos.reshape(-1, 1).astype(np.float64) * freq[None, :]
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def sincos_pos_embed_3d(embed_dim: int, grid_size: int, cls_token: bool = True) -> np.ndarray:
    ar = np.arange(grid_size, dtype=np.float32)
    a, b, c = np.meshgrid(ar, ar, ar)          # 'xy' indexing, as in the reference
    per_axis = embed_dim // 3
    if per_axis % 2:
        per_axis += 1
    rest = embed_dim - 2 * per_axis
    table = np.concatenate([_sincos_1d(per_axis, a), _sincos_1d(per_axis, b), _sincos_1d(rest, c)], axis=1)
    if cls_token:
        table = np.concatenate([np.zeros((1, embed_dim)), table], axis=0)
    return table


class EdgeMapLoss(torch.nn.Module):
    """mse(sobel(pred_vol), sobel(gaussian_blur(target_vol, sigma=2))) -- model/vit_autoenc.py:221-224.

    Sobel: model/model_utils/sobel_filter.py:10-45 (3 directional 3x3x3 kernels per channel, sqrt of the squared sum,
    summed over channels).  Blur: model/model_utils/gaussian_filter.py:5-26 (ks = int(5*sigma) rounded up to odd = 11,
    taps at linspace(-ks//2, ks//2+1, ks) -- i.e. 1.2 apart -- normalised, zero padding); applied here as three 1-D
    passes, which equals the reference's dense ks^3 kernel (outer product of the same taps) under zero padding."""

    def __init__(self, sigma: float = 2.0):
        super().__init__()
        s = torch.tensor([1.0, 2.0, 1.0])
        d = torch.tensor([1.0, 0.0, -1.0])
        k = torch.stack([torch.einsum("i,j,k->ijk", s, s, d), torch.einsum("i,j,k->ijk", s, -d, s),
                         torch.einsum("i,j,k->ijk", -d, s, s)]).unsqueeze(1)
        self.register_buffer("sobel", k, persistent=False)
        ks = int(sigma * 5)
        ks += 1 - ks % 2
        ts = torch.linspace(-ks // 2, ks // 2 + 1, ks)
        taps = torch.exp(-(ts / sigma) ** 2 / 2)
        self.register_buffer("taps", taps / taps.sum(), persistent=False)

    def edge_map(self, vol: torch.Tensor) -> torch.Tensor:
        B, C = vol.shape[:2]
        g = F.conv3d(vol.reshape(B * C, 1, *vol.shape[2:]), self.sobel, padding=1)
        return torch.sqrt((g * g).sum(dim=1)).reshape(B, C, *vol.shape[2:]).sum(dim=1)

    def blur(self, vol: torch.Tensor) -> torch.Tensor:
        B, C = vol.shape[:2]
        x = vol.reshape(B * C, 1, *vol.shape[2:])
        n = self.taps.numel()
        x = F.conv3d(x, self.taps.view(1, 1, n, 1, 1), padding=(n // 2, 0, 0))
        x = F.conv3d(x, self.taps.view(1, 1, 1, n, 1), padding=(0, n // 2, 0))
        x = F.conv3d(x, self.taps.view(1, 1, 1, 1, n), padding=(0, 0, n // 2))
        return x.reshape(vol.shape)

    def forward(self, pred_vol: torch.Tensor, target_vol: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            tgt = self.edge_map(self.blur(target_vol))
        return F.mse_loss(self.edge_map(pred_vol), tgt, reduction="mean")
