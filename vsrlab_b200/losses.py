"""Fused training objective and on-device metrics for the callers of the hot path.

The reference computes `CharbonnierLoss(sr, hr) + CharbonnierLoss(lq, resize(hr, (h, w)))` with a dozen elementwise
torch kernels and then synchronises the host twice per step for `.item()` of PSNR / SSIM (train.py:93-101,
core/utils.py:235-252, core/metrics.py:13).  Here each loss term is ONE kernel forward (read x, y; reduce) and ONE
kernel backward (read x, y; write the gradient, scaled by the upstream gradient taken from device memory), the resize
of `hr` is computed inside the second term's kernel, and the metrics return device tensors.

* `CharbonnierLoss`          drop-in for `vsrlab.core.losses.CharbonnierLoss` (same constructor, same forward(x, y))
* `realbasicvsr_loss`        both terms of `compute_loss` (core/utils.py:235-240) in two launches
* `PSNR`, `SSIM`             modules with piqa's call signature (defaults of piqa.PSNR / piqa.SSIM), results stay on
                             the device; `MetricCollection.forward` of the reference calls `.item()` on them, which works.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from .ops import _p, _stream, require_cuda


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


class _CharbonnierFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, eps):
        require_cuda(x, "charbonnier input")
        xc, yc = _f32c(x), _f32c(y)
        if xc.shape != yc.shape:
            raise L.VsrbError(f"CharbonnierLoss: shapes differ {tuple(xc.shape)} vs {tuple(yc.shape)}")
        acc = torch.zeros(1, dtype=torch.float64, device=x.device)
        L.check(L.load().vsrb_charbonnier(_p(xc), _p(yc), xc.numel(), eps, _p(acc), None, None, 1.0, _stream()), "vsrb_charbonnier")
        ctx.save_for_backward(xc, yc)
        ctx.eps, ctx.x_dtype, ctx.y_dtype = eps, x.dtype, y.dtype
        return (acc / xc.numel()).to(torch.float32).squeeze(0)

    @staticmethod
    def backward(ctx, g):
        xc, yc = ctx.saved_tensors
        gx = torch.empty_like(xc)
        acc = torch.zeros(1, dtype=torch.float64, device=xc.device)
        gs = g.detach().to(torch.float32).contiguous()
        L.check(L.load().vsrb_charbonnier(_p(xc), _p(yc), xc.numel(), ctx.eps, _p(acc), _p(gx), _p(gs), 1.0 / xc.numel(), _stream()),
                "vsrb_charbonnier")
        gy = -gx if ctx.needs_input_grad[1] else None
        return (gx.to(ctx.x_dtype) if ctx.needs_input_grad[0] else None), (gy.to(ctx.y_dtype) if gy is not None else None), None


class _CharbonnierResizedFn(torch.autograd.Function):
    """mean(sqrt((lq - resize(hr, lq.shape[-2:]))^2 + eps)); gradient wrt lq only (hr is data)."""

    @staticmethod
    def forward(ctx, lq, hr, eps):
        require_cuda(lq, "charbonnier input")
        lc, hc = _f32c(lq), _f32c(hr)
        h, w = lc.shape[-2:]
        H, W = hc.shape[-2:]
        planes = lc.numel() // (h * w)
        if hc.numel() // (H * W) != planes:
            raise L.VsrbError("charbonnier_resized: lq and hr must have the same leading dimensions")
        acc = torch.zeros(1, dtype=torch.float64, device=lq.device)
        L.check(L.load().vsrb_charbonnier_resized(_p(lc), _p(hc), planes, h, w, H, W, eps, _p(acc), None, None, 1.0, _stream()),
                "vsrb_charbonnier_resized")
        ctx.save_for_backward(lc, hc)
        ctx.eps, ctx.dtype = eps, lq.dtype
        return (acc / lc.numel()).to(torch.float32).squeeze(0)

    @staticmethod
    def backward(ctx, g):
        lc, hc = ctx.saved_tensors
        h, w = lc.shape[-2:]
        H, W = hc.shape[-2:]
        planes = lc.numel() // (h * w)
        gl = torch.empty_like(lc)
        acc = torch.zeros(1, dtype=torch.float64, device=lc.device)
        gs = g.detach().to(torch.float32).contiguous()
        L.check(L.load().vsrb_charbonnier_resized(_p(lc), _p(hc), planes, h, w, H, W, ctx.eps, _p(acc), _p(gl), _p(gs), 1.0 / lc.numel(),
                                                  _stream()), "vsrb_charbonnier_resized")
        return gl.to(ctx.dtype), None, None


class CharbonnierLoss(torch.nn.Module):
    """Drop-in for the reference's CharbonnierLoss (core/losses.py:10-18): forward(x, y) = mean(sqrt((x-y)^2 + eps))."""

    def __init__(self, eps: float = 1e-9):
        super().__init__()
        self.eps = eps

    def forward(self, x, y):
        return _CharbonnierFn.apply(x, y, self.eps)


def realbasicvsr_loss(sr: torch.Tensor, hr: torch.Tensor, lq: Optional[torch.Tensor] = None, eps: float = 1e-9) -> torch.Tensor:
    """compute_loss of the reference (core/utils.py:235-240) with its Charbonnier loss: the resize of `hr` to lq's
    size happens inside the second term's kernel instead of materialising the resized clip."""
    loss = _CharbonnierFn.apply(sr, hr, eps)
    if lq is not None:
        loss = loss + _CharbonnierResizedFn.apply(lq, hr, eps)
    return loss


class PSNR(torch.nn.Module):
    """piqa.PSNR semantics (epsilon 1e-8, value_range 1, reduction 'mean') on [N,C,H,W]; `x` is clamped to [0,1] inside the
    kernel (what core/utils.py:244 does before the call - clamping twice changes nothing)."""

    def __init__(self, epsilon: float = 1e-8, value_range: float = 1.0, reduction: str = "mean"):
        super().__init__()
        if value_range != 1.0:
            raise L.VsrbError("PSNR: value_range 1 only (frames are [0,1]-scaled on this path)")
        self.epsilon, self.reduction = epsilon, reduction

    def forward(self, x, y):
        require_cuda(x, "metric input")
        xc, yc = _f32c(x), _f32c(y)
        n = xc.shape[0]
        per = xc.numel() // n
        sums = torch.zeros(n, dtype=torch.float64, device=x.device)
        L.check(L.load().vsrb_psnr_sums(_p(xc), _p(yc), n, per, _p(sums), _stream()), "vsrb_psnr_sums")
        v = (10.0 * torch.log10(1.0 / (sums / per + self.epsilon))).to(torch.float32)
        return v.mean() if self.reduction == "mean" else (v.sum() if self.reduction == "sum" else v)


class SSIM(torch.nn.Module):
    """piqa.SSIM defaults (window 11, sigma 1.5, k1 0.01, k2 0.03, value_range 1, per-channel Gaussian filter with 'valid'
    borders, reduction 'mean') on [N,C,H,W], H, W >= 11."""

    def __init__(self, window_size: int = 11, sigma: float = 1.5, n_channels: int = 3, reduction: str = "mean", **kwargs):
        super().__init__()
        if window_size != 11 or sigma != 1.5 or kwargs.get("value_range", 1.0) != 1.0:
            raise L.VsrbError("SSIM: the fused kernel implements piqa's defaults (window 11, sigma 1.5, value_range 1)")
        self.reduction = reduction

    def forward(self, x, y):
        require_cuda(x, "metric input")
        xc, yc = _f32c(x), _f32c(y)
        n, c, h, w = xc.shape
        sums = torch.zeros(n, dtype=torch.float64, device=x.device)
        L.check(L.load().vsrb_ssim_sums(_p(xc), _p(yc), n, c, h, w, 1, _p(sums), _stream()), "vsrb_ssim_sums")
        v = (sums / (c * (h - 10) * (w - 10))).to(torch.float32)
        return v.mean() if self.reduction == "mean" else (v.sum() if self.reduction == "sum" else v)


def running_metrics_on_device(sr: torch.Tensor, hr: torch.Tensor):
    """compute_metric of the reference (core/utils.py:242-247) without host synchronisation: {'PSNR': t, 'SSIM': t}."""
    x = sr.detach().flatten(0, 1) if sr.dim() == 5 else sr.detach()
    y = hr.detach().flatten(0, 1) if hr.dim() == 5 else hr.detach()
    return {"PSNR": PSNR()(x, y), "SSIM": SSIM()(x, y)}
