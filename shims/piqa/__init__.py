"""Minimal `piqa`: PSNR and SSIM modules with piqa's defaults, the two metrics the reference's configs instantiate
(conf/train/default.yaml:9-15, conf/experiment/test.yaml:10-16).  On CUDA tensors they run the fused kernels of
`vsrlab_b200.losses` (one launch each, no host synchronisation); on CPU tensors a plain torch restatement."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _gaussian(window: int, sigma: float, dtype, device):
    k = torch.arange(window, dtype=torch.float64) - (window - 1) / 2
    g = torch.exp(-k ** 2 / (2 * sigma ** 2))
    return (g / g.sum()).to(dtype).to(device)


class PSNR(torch.nn.Module):
    def __init__(self, epsilon: float = 1e-8, value_range: float = 1.0, reduction: str = "mean"):
        super().__init__()
        self.epsilon, self.value_range, self.reduction = epsilon, value_range, reduction

    def forward(self, x, y):
        if x.is_cuda and self.value_range == 1.0:
            from vsrlab_b200.losses import PSNR as _Fused
            return _Fused(self.epsilon, 1.0, self.reduction)(x, y)
        mse = ((x - y) ** 2).flatten(1).mean(-1)
        v = 10 * torch.log10(self.value_range ** 2 / (mse + self.epsilon))
        return v.mean() if self.reduction == "mean" else (v.sum() if self.reduction == "sum" else v)


class SSIM(torch.nn.Module):
    def __init__(self, window_size: int = 11, sigma: float = 1.5, n_channels: int = 3, reduction: str = "mean", value_range: float = 1.0,
                 k1: float = 0.01, k2: float = 0.03):
        super().__init__()
        self.window_size, self.sigma, self.reduction, self.value_range, self.k1, self.k2 = window_size, sigma, reduction, value_range, k1, k2

    def forward(self, x, y):
        default = (self.window_size, self.sigma, self.value_range, self.k1, self.k2) == (11, 1.5, 1.0, 0.01, 0.03)
        if x.is_cuda and default:
            from vsrlab_b200.losses import SSIM as _Fused
            return _Fused(reduction=self.reduction)(x, y)
        c = x.shape[1]
        g = _gaussian(self.window_size, self.sigma, x.dtype, x.device)
        win = (g[:, None] * g[None, :]).expand(c, 1, self.window_size, self.window_size).contiguous()

        def filt(t):
            return F.conv2d(t, win, groups=c)
        mx, my = filt(x), filt(y)
        mxx, myy, mxy = mx * mx, my * my, mx * my
        sxx, syy, sxy = filt(x * x) - mxx, filt(y * y) - myy, filt(x * y) - mxy
        c1, c2 = (self.k1 * self.value_range) ** 2, (self.k2 * self.value_range) ** 2
        ss = (2 * mxy + c1) / (mxx + myy + c1) * ((2 * sxy + c2) / (sxx + syy + c2))
        v = ss.flatten(1).mean(-1)
        return v.mean() if self.reduction == "mean" else (v.sum() if self.reduction == "sum" else v)
