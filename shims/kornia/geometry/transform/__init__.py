"""`kornia.geometry.transform.resize` as the reference uses it (core/utils.py:239, core/loggers.py:39,44,
vsr/dataset.py:54): bilinear, align_corners=None (-> False), no antialias, any number of leading dimensions."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def resize(input: torch.Tensor, size, interpolation: str = "bilinear", align_corners=None, side: str = "short",
           antialias: bool = False) -> torch.Tensor:
    if isinstance(size, int):
        h, w = input.shape[-2:]
        if (side == "short") == (h <= w):
            size = (size, max(1, round(w * size / h)))
        else:
            size = (max(1, round(h * size / w)), size)
    lead = input.shape[:-3]
    x = input.reshape(-1, *input.shape[-3:]) if input.dim() != 4 else input
    if tuple(x.shape[-2:]) == tuple(size):
        out = x
    else:
        kw = {} if interpolation in ("nearest", "area") else {"align_corners": bool(align_corners) if align_corners is not None else False}
        out = F.interpolate(x, size=tuple(size), mode=interpolation, antialias=antialias, **kw)
    return out.reshape(*lead, *out.shape[-3:]) if input.dim() != 4 else out
