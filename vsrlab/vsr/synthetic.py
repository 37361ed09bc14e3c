"""Synthetic video clips with the interface of the reference's `DatasetVSR` (src/vsr/dataset.py:16-65): item ->
(lr_video [T,3,h,w], hr_video [T,3,H,W]) float32 in [0,1], LR = bilinear downscale of HR by `scale` as at
vsr/dataset.py:52-54.  Exists because the reference's data group `conf/train/data/*` is not in its repository and no
dataset can be fetched in this environment; selected by `conf/train/data/{default,gan}.yaml` of this repository."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch.utils.data import Dataset


class SyntheticVSR(Dataset):
    def __init__(self, split: str = "train", length: int = 256, seq: int = 15, scale: int = 4, lr_size=(64, 64), seed: int = 0):
        super().__init__()
        self.split, self.length, self.seq, self.scale, self.seed = split, int(length), int(seq), int(scale), int(seed)
        self.lr_size = tuple(int(v) for v in lr_size)

    def __len__(self) -> int:
        return self.length

    def __getitem__(self, index: int):
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + index)
        h, w = self.lr_size
        H, W = h * self.scale, w * self.scale
        # a smooth random texture translating with a constant sub-pixel velocity: frames are related by real motion,
        # so flow estimation and temporal propagation see something meaningful
        pad = 16
        base = F.interpolate(torch.rand(1, 3, (H + 2 * pad) // 8 + 2, (W + 2 * pad) // 8 + 2, generator=g), size=(H + 2 * pad, W + 2 * pad),
                             mode="bicubic", align_corners=False).clamp(0, 1)
        fine = torch.rand(1, 3, H + 2 * pad, W + 2 * pad, generator=g) * 0.08
        tex = (base * 0.92 + fine)[0]
        ang = torch.rand(1, generator=g).item() * 2 * math.pi
        speed = torch.rand(1, generator=g).item() * 0.9
        frames = []
        ys = torch.arange(H, dtype=torch.float32).view(H, 1).expand(H, W)
        xs = torch.arange(W, dtype=torch.float32).view(1, W).expand(H, W)
        for t in range(self.seq):
            dx, dy = speed * t * math.cos(ang), speed * t * math.sin(ang)
            gx = (xs + pad + dx) / (W + 2 * pad - 1) * 2 - 1
            gy = (ys + pad + dy) / (H + 2 * pad - 1) * 2 - 1
            frames.append(F.grid_sample(tex[None], torch.stack([gx, gy], -1)[None], mode="bilinear", align_corners=True)[0])
        hr = torch.stack(frames).clamp(0, 1)
        lr = F.interpolate(hr, size=(h, w), mode="bilinear", align_corners=False)
        return lr, hr
