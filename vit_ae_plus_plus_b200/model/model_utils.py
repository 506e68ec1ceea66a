"""Init-time host helper of the MAE path: the frozen 3-D sin-cos positional table.

* ``sincos_pos_embed_3d`` follows the reference's ``get_3d_sincos_pos_embed`` (model/model_utils/vit_helpers.py:13-70),
  including its quirks: ``np.meshgrid`` default 'xy' indexing (so the first channel third encodes the H index, the
  second the D index, the last the W index), channel split ``res = ceil_even(D // 3)`` with the remainder on the last
  axis, and an all-zero cls row.  fp64 numpy, cast to fp32 by the caller -- init only, never on the step.
"""
from __future__ import annotations

import numpy as np


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
