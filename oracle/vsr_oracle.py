"""CPU oracle for the Real-BasicVSR / BasicVSR hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* of the algorithm the reference implements with
stock PyTorch modules (santurini/vsrlab, `src/vsr/models/RealBasicVSR/**`,
`src/core/modules/{conv,upsampling}.py`).  It exists so that the CUDA path in
`vsrlab_b200/` can be checked bit-for-tolerance against something that runs on
a CPU; it is never imported by the product path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.

Parity pin: every function here is checked against the outputs of the
reference's own modules (imported live from /root/reference by
`tests/golden/make_golden.py`, outputs committed under `tests/golden/*.npz`)
in `tests/test_oracle_golden.py`.  The reference ships no tests or golden
vectors of its own, so those live-import fixtures are the pin.

Style: purely functional, parameters come in as a flat ``state_dict``-style
mapping with the reference's key names.  All resampling (bilinear resize,
avg-pool, backward warp, pixel shuffle) is written from first principles with
explicit index arithmetic so that it states the semantics the CUDA kernels
implement, instead of calling the same library entry points the reference
calls.  The dense contraction itself is ``F.conv2d`` (fp32, CPU).
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Tuple

import torch
import torch.nn.functional as F

Params = Mapping[str, torch.Tensor]

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # reference spynet.py:30
IMAGENET_STD = (0.229, 0.224, 0.225)    # reference spynet.py:31


# --------------------------------------------------------------------------
# dense contraction + pointwise
# --------------------------------------------------------------------------
def _k(prefix: str, name: str) -> str:
    return f"{prefix}.{name}" if prefix else name


def conv(x: torch.Tensor, P: Params, prefix: str) -> torch.Tensor:
    """Stride-1 'same' convolution with bias (every conv on the path:
    conv.py:85-86,97; upsampling.py:7; basicvsr.py:18,20-21; realbasicvsr.py:22;
    spynet.py:16-18)."""
    w = P[_k(prefix, "weight")]
    b = P[_k(prefix, "bias")]
    return F.conv2d(x, w, b, stride=1, padding=w.shape[-1] // 2)


def lrelu(x: torch.Tensor, slope: float = 0.1) -> torch.Tensor:
    return torch.where(x >= 0, x, x * slope)


def relu(x: torch.Tensor) -> torch.Tensor:
    return torch.clamp_min(x, 0.0)


def count_blocks(P: Params, prefix: str) -> int:
    n = 0
    while _k(prefix, f"res_block.{n}.conv1.weight") in P:
        n += 1
    return n


def residual_block(x: torch.Tensor, P: Params, prefix: str) -> torch.Tensor:
    """ResidualBlock: stem conv + LeakyReLU(0.1), then B x (conv, ReLU, conv, +id).
    Reference conv.py:94-103 (stack) and conv.py:82-92 (one ResidualConv)."""
    x = lrelu(conv(x, P, _k(prefix, "conv.0")))
    for k in range(count_blocks(P, prefix)):
        t = relu(conv(x, P, _k(prefix, f"res_block.{k}.conv1")))
        x = conv(t, P, _k(prefix, f"res_block.{k}.conv2")) + x
    return x


# --------------------------------------------------------------------------
# resampling, first principles
# --------------------------------------------------------------------------
def _src_index(out_size: int, in_size: int, align_corners: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Source taps and weight for 1-D linear interpolation, the way ATen's
    upsample_bilinear2d defines them (area_pixel_compute_source_index)."""
    o = torch.arange(out_size, dtype=torch.float32)
    if align_corners:
        scale = (in_size - 1) / (out_size - 1) if out_size > 1 else 0.0
        src = o * torch.tensor(scale, dtype=torch.float32)
    else:
        scale = in_size / out_size
        src = (o + 0.5) * torch.tensor(scale, dtype=torch.float32) - 0.5
        src = torch.clamp_min(src, 0.0)
    i0 = src.floor().to(torch.int64).clamp_max(in_size - 1)
    i1 = (i0 + 1).clamp_max(in_size - 1)
    lam = (src - i0.to(torch.float32)).clamp(0.0, 1.0)
    return i0, i1, lam


def bilinear_resize(x: torch.Tensor, size: Tuple[int, int], align_corners: bool) -> torch.Tensor:
    """F.interpolate(mode='bilinear') restated (spynet.py:54,74-87; basicvsr.py:22)."""
    H, W = size
    y0, y1, ly = _src_index(H, x.shape[-2], align_corners)
    x0, x1, lx = _src_index(W, x.shape[-1], align_corners)
    top = x[..., y0, :]
    bot = x[..., y1, :]
    ly = ly.view(-1, 1)
    rows = top * (1 - ly) + bot * ly
    return rows[..., x0] * (1 - lx) + rows[..., x1] * lx


def avg_pool2(x: torch.Tensor) -> torch.Tensor:
    """avg_pool2d(kernel 2, stride 2) on even sizes (spynet.py:44-45)."""
    return (x[..., 0::2, 0::2] + x[..., 0::2, 1::2] + x[..., 1::2, 0::2] + x[..., 1::2, 1::2]) * 0.25


def pixel_shuffle(x: torch.Tensor, r: int) -> torch.Tensor:
    """out[n,c,r*y+i,r*x+j] = x[n, c*r*r + i*r + j, y, x] (upsampling.py:8,12)."""
    n, c, h, w = x.shape
    co = c // (r * r)
    return x.view(n, co, r, r, h, w).permute(0, 1, 4, 2, 5, 3).reshape(n, co, h * r, w * r)


def flow_warp(x: torch.Tensor, flow: torch.Tensor, padding_mode: str = "zeros") -> torch.Tensor:
    """Backward warp: out[n,c,y,x] = bilinear(x[n,c], (x+flow[...,0], y+flow[...,1])).

    Reference spynet.py:95-106: the pixel coordinate is normalised to [-1,1]
    (``2*g/max(size-1,1)-1``) and handed to grid_sample(align_corners=True), which
    maps it back with ``(g+1)/2*(size-1)``.  Both roundings are reproduced so
    that the fp32 sample position is the same.  'zeros': taps outside the image
    contribute 0.  'border': the sample position is clamped to [0, size-1].
    ``flow`` is channels-last [N,h,w,2] exactly as the reference callers pass it
    (basicvsr.py:54,69; spynet.py:59).
    """
    n, c, h, w = x.shape
    gy, gx = torch.meshgrid(torch.arange(h, dtype=x.dtype), torch.arange(w, dtype=x.dtype), indexing="ij")
    px = gx + flow[..., 0]
    py = gy + flow[..., 1]
    nx = 2.0 * px / max(w - 1, 1) - 1.0
    ny = 2.0 * py / max(h - 1, 1) - 1.0
    ix = (nx + 1.0) / 2.0 * (w - 1)
    iy = (ny + 1.0) / 2.0 * (h - 1)
    if padding_mode == "border":
        ix = ix.clamp(0.0, float(w - 1))
        iy = iy.clamp(0.0, float(h - 1))
    x0f = ix.floor()
    y0f = iy.floor()
    wx1 = ix - x0f
    wy1 = iy - y0f
    wx0 = 1.0 - wx1
    wy0 = 1.0 - wy1
    x0 = x0f.to(torch.int64)
    y0 = y0f.to(torch.int64)
    flat = x.reshape(n, c, h * w)
    out = torch.zeros_like(x)
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xi = x0 + dx
            yi = y0 + dy
            ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
            idx = (yi.clamp(0, h - 1) * w + xi.clamp(0, w - 1)).view(n, 1, h * w).expand(n, c, h * w)
            tap = torch.gather(flat, 2, idx).view(n, c, h, w)
            out = out + tap * (wy * wx * ok.to(x.dtype)).unsqueeze(1)
    return out


# --------------------------------------------------------------------------
# SPyNet (reference modules/spynet.py)
# --------------------------------------------------------------------------
def spynet_module(x: torch.Tensor, P: Params, prefix: str) -> torch.Tensor:
    """Five 7x7 convs, ReLU after each including the last (spynet.py:13-21)."""
    for j in range(5):
        x = relu(conv(x, P, _k(prefix, f"basic_module.{j}.conv.0")))
    return x


def spynet_compute_flow(ref: torch.Tensor, supp: torch.Tensor, P: Params, prefix: str) -> torch.Tensor:
    """Coarse-to-fine flow on sizes that are multiples of 32 (spynet.py:38-67)."""
    mean = P[_k(prefix, "mean")] if _k(prefix, "mean") in P else torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = P[_k(prefix, "std")] if _k(prefix, "std") in P else torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    t, _, h, w = ref.shape
    refs = [(ref - mean) / std]
    supps = [(supp - mean) / std]
    for _ in range(5):
        refs.append(avg_pool2(refs[-1]))
        supps.append(avg_pool2(supps[-1]))
    refs, supps = refs[::-1], supps[::-1]
    flow = ref.new_zeros(t, 2, h // 32, w // 32)
    for level in range(6):
        if level == 0:
            flow_up = flow
        else:
            flow_up = bilinear_resize(flow, (flow.shape[-2] * 2, flow.shape[-1] * 2), align_corners=True) * 2.0
        warped = flow_warp(supps[level], flow_up.permute(0, 2, 3, 1), padding_mode="border")
        residue = spynet_module(torch.cat([refs[level], warped, flow_up], 1), P, _k(prefix, f"basic_module.{level}"))
        flow = flow_up + residue
    return flow


def spynet(ref: torch.Tensor, supp: torch.Tensor, P: Params, prefix: str = "") -> torch.Tensor:
    """Resize to /32, estimate, resize back, rescale (spynet.py:69-93)."""
    h, w = ref.shape[-2:]
    w_up = w if w % 32 == 0 else 32 * (w // 32 + 1)
    h_up = h if h % 32 == 0 else 32 * (h // 32 + 1)
    ref = bilinear_resize(ref, (h_up, w_up), align_corners=False)
    supp = bilinear_resize(supp, (h_up, w_up), align_corners=False)
    flow = bilinear_resize(spynet_compute_flow(ref, supp, P, prefix.rstrip(".")), (h, w), align_corners=False)
    scale = torch.tensor([float(w) / float(w_up), float(h) / float(h_up)], dtype=flow.dtype).view(1, 2, 1, 1)
    return flow * scale


# --------------------------------------------------------------------------
# BasicVSR / Real-BasicVSR
# --------------------------------------------------------------------------
def basicvsr_flows(lrs: torch.Tensor, P: Params, prefix: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """(flow_forward, flow_backward), each [n, t-1, 2, h, w] (basicvsr.py:30-43)."""
    n, t, c, h, w = lrs.shape
    a = lrs[:, :-1].reshape(-1, c, h, w)
    b = lrs[:, 1:].reshape(-1, c, h, w)
    fb = spynet(a, b, P, prefix + ".spynet")
    ff = spynet(b, a, P, prefix + ".spynet")
    return ff.view(n, t - 1, 2, h, w), fb.view(n, t - 1, 2, h, w)


def basicvsr(lrs: torch.Tensor, P: Params, prefix: str = "basicvsr", return_flows: bool = False):
    """Bidirectional recurrent propagation + reconstruction (basicvsr.py:39-83)."""
    n, t, c, h, w = lrs.shape
    mid = P[prefix + ".point_conv.0.weight"].shape[0]
    ff, fb = basicvsr_flows(lrs, P, prefix)

    back = [None] * t
    feat = lrs.new_zeros(n, mid, h, w)
    for i in range(t - 1, -1, -1):
        if i < t - 1:
            feat = flow_warp(feat, fb[:, i].permute(0, 2, 3, 1))
        feat = residual_block(torch.cat([lrs[:, i], feat], 1), P, prefix + ".backward_resblocks")
        back[i] = feat

    n_up = 0
    while f"{prefix}.upsample.{n_up}.upconv.weight" in P:
        n_up += 1
    scale = 2 ** n_up
    outs = []
    feat = torch.zeros_like(feat)
    for i in range(t):
        if i > 0:
            feat = flow_warp(feat, ff[:, i - 1].permute(0, 2, 3, 1))
        feat = residual_block(torch.cat([lrs[:, i], feat], 1), P, prefix + ".forward_resblocks")
        out = lrelu(conv(torch.cat([back[i], feat], 1), P, prefix + ".point_conv.0"))
        for u in range(n_up):                      # no activation between packs
            out = pixel_shuffle(conv(out, P, f"{prefix}.upsample.{u}.upconv"), 2)
        out = conv(lrelu(conv(out, P, prefix + ".conv_last.0")), P, prefix + ".conv_last.2")
        outs.append(out + bilinear_resize(lrs[:, i], (h * scale, w * scale), align_corners=False))
    sr = torch.stack(outs, 1)
    return (sr, ff, fb) if return_flows else sr


def cleaner(lr: torch.Tensor, P: Params, prefix: str = "cleaner", steps: int = 3) -> torch.Tensor:
    """IterativeRefinement: x += conv(resblock(x)), fixed 3 rounds (realbasicvsr.py:17-30).
    Returns a new tensor; the in-place/aliasing contract of the reference is a
    property of the host mirror and is tested there."""
    n, t, c, h, w = lr.shape
    x = lr.reshape(n * t, c, h, w).clone()
    for _ in range(steps):
        x = x + conv(residual_block(x, P, prefix + ".resblock"), P, prefix + ".conv")
    return x.view(n, t, c, h, w)


def realbasicvsr(lr: torch.Tensor, P: Params, return_flows: bool = False):
    """(sr, lq) = RealBasicVSR.forward (realbasicvsr.py:11-15)."""
    lq = cleaner(lr, P)
    out = basicvsr(lq, P, "basicvsr", return_flows=return_flows)
    if return_flows:
        return out[0], lq, out[1], out[2]
    return out, lq


def pixel_shuffle_pack(x: torch.Tensor, P: Params, prefix: str, r: int = 2) -> torch.Tensor:
    """PixelShufflePack.forward (upsampling.py:10-12)."""
    return pixel_shuffle(conv(x, P, _k(prefix, "upconv")), r)


# --------------------------------------------------------------------------
# helpers shared by tests / bench (not part of the algorithm)
# --------------------------------------------------------------------------
def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    mse = torch.mean((a.clamp(0, 1).double() - b.clamp(0, 1).double()) ** 2).item()
    return 99.0 if mse == 0 else 10.0 * math.log10(1.0 / mse)


def conv_flops(n: int, t: int, h: int, w: int, cleaning_blocks: int, res_blocks: int) -> Dict[str, float]:
    """Algorithmic conv FLOPs of one forward (SURVEY.md §8d formulae)."""
    px = n * t * h * w
    hp = -(-h // 32) * 32
    wp = -(-w // 32) * 32
    lvl_px = sum((hp >> k) * (wp >> k) for k in range(6))
    d = {
        "cleaner": 3 * (2 * 27 * 64 + 2 * 576 * 3 + 147456 * cleaning_blocks) * px,
        "prop": 2 * (2 * 67 * 9 * 64 + 147456 * res_blocks) * px,
        "point": 2 * 128 * 64 * px,
        "upsample": (2 * 576 * 256) * px * (1 + 4),
        "last": (2 * 576 * 64 + 2 * 576 * 3) * px * 16,
        "spynet": 479808.0 * lvl_px * 2 * n * (t - 1),
    }
    d["total"] = float(sum(d.values()))
    return d
