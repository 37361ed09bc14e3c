"""Fused training objective and on-device metrics against plain torch restatements of what the reference
computes: CharbonnierLoss (core/losses.py:10-18), compute_loss with kornia's bilinear resize (core/utils.py:235-240), and
piqa.PSNR / piqa.SSIM with their default arguments (conf/train/default.yaml:9-15).  piqa and kornia are not installable
here (no network), so their published algorithms are restated below; that part of the parity is therefore unpinned."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from vsrlab_b200 import load
    load()
    return torch.device("cuda:0")


def charbonnier_ref(x, y, eps=1e-9):
    d = x - y
    return torch.mean(torch.sqrt(d * d + eps))


def ssim_ref(x, y):
    """piqa.SSIM defaults: per-channel 11x11 Gaussian (sigma 1.5) 'valid' filtering, k1 0.01, k2 0.03, L = 1."""
    k = torch.arange(11, dtype=torch.float64) - 5
    g = torch.exp(-k ** 2 / (2 * 1.5 ** 2))
    g = (g / g.sum()).to(x.dtype).to(x.device)
    c = x.shape[1]
    win = (g[:, None] * g[None, :]).expand(c, 1, 11, 11).contiguous()

    def filt(t):
        return F.conv2d(t, win, groups=c)
    mx, my = filt(x), filt(y)
    mxx, myy, mxy = mx * mx, my * my, mx * my
    sxx, syy, sxy = filt(x * x) - mxx, filt(y * y) - myy, filt(x * y) - mxy
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    cs = (2 * sxy + c2) / (sxx + syy + c2)
    ss = (2 * mxy + c1) / (mxx + myy + c1) * cs
    return ss.flatten(1).mean(-1)


def psnr_ref(x, y, eps=1e-8):
    mse = ((x - y) ** 2).flatten(1).mean(-1)
    return 10 * torch.log10(1.0 / (mse + eps))


@pytest.mark.parametrize("shape", [(2, 3, 3, 64, 64), (1, 2, 3, 37, 53), (7,)])
def test_charbonnier_matches_torch_forward_and_backward(dev, shape):
    from vsrlab_b200.losses import CharbonnierLoss
    g = torch.Generator().manual_seed(1)
    x = torch.rand(*shape, generator=g).to(dev).requires_grad_(True)
    y = torch.rand(*shape, generator=g).to(dev)
    xr = x.detach().clone().double().requires_grad_(True)
    ref = charbonnier_ref(xr, y.double())
    (ref * 3.0).backward()
    loss = CharbonnierLoss()(x, y)
    assert loss.dtype == torch.float32 and loss.dim() == 0
    (loss * 3.0).backward()
    assert abs(loss.item() - ref.item()) <= 1e-6 * max(1.0, abs(ref.item()))
    assert (x.grad.double() - xr.grad).abs().max().item() <= 1e-5 * xr.grad.abs().max().item() + 1e-12


def test_realbasicvsr_loss_matches_compute_loss(dev):
    """Both terms of core/utils.py:235-240 incl. the bilinear (align_corners=False) resize of hr, and GradScaler-style
    scaling of the upstream gradient taken from device memory."""
    from vsrlab_b200.losses import realbasicvsr_loss
    g = torch.Generator().manual_seed(2)
    sr = torch.rand(2, 3, 3, 96, 128, generator=g).to(dev).requires_grad_(True)
    hr = torch.rand(2, 3, 3, 96, 128, generator=g).to(dev)
    lq = torch.rand(2, 3, 3, 24, 32, generator=g).to(dev).requires_grad_(True)
    sr_r, lq_r = sr.detach().clone().requires_grad_(True), lq.detach().clone().requires_grad_(True)
    ref = charbonnier_ref(sr_r, hr) + charbonnier_ref(lq_r, F.interpolate(hr.flatten(0, 1), size=(24, 32), mode="bilinear",
                                                                           align_corners=False).view_as(lq_r))
    (ref * 65536.0 / 4).backward()
    loss = realbasicvsr_loss(sr, hr, lq)
    (loss * 65536.0 / 4).backward()
    assert abs(loss.item() - ref.item()) <= 2e-6
    for a, b in ((sr.grad, sr_r.grad), (lq.grad, lq_r.grad)):
        assert (a - b).abs().max().item() <= 2e-5 * b.abs().max().item()
    # non-integer scale factor (kornia resize to an arbitrary size)
    lq2 = torch.rand(1, 2, 3, 25, 37, generator=g).to(dev)
    hr2 = torch.rand(1, 2, 3, 96, 128, generator=g).to(dev)
    want = charbonnier_ref(lq2, F.interpolate(hr2.flatten(0, 1), size=(25, 37), mode="bilinear", align_corners=False).view_as(lq2))
    from vsrlab_b200.losses import _CharbonnierResizedFn
    assert abs(_CharbonnierResizedFn.apply(lq2, hr2, 1e-9).item() - want.item()) <= 2e-6


@pytest.mark.parametrize("shape", [(6, 3, 64, 64), (2, 3, 37, 53), (1, 3, 11, 11), (3, 1, 720, 1280)])
def test_psnr_ssim_match_piqa_formulas(dev, shape):
    from vsrlab_b200.losses import PSNR, SSIM
    g = torch.Generator().manual_seed(3)
    y = torch.rand(*shape, generator=g).to(dev)
    x = (y + 0.1 * torch.randn(*shape, generator=g).to(dev))               # leaves [0,1] in places: the clamp matters
    xc = x.clamp(0, 1)
    p, s = PSNR()(x, y), SSIM(n_channels=shape[1])(x, y)
    assert p.is_cuda and s.is_cuda and p.dim() == 0
    assert abs(p.item() - psnr_ref(xc.double(), y.double()).mean().item()) <= 1e-4
    assert abs(s.item() - ssim_ref(xc.double(), y.double()).mean().item()) <= 2e-5
    per = SSIM(reduction="none")(x, y)
    assert (per.double() - ssim_ref(xc.double(), y.double())).abs().max().item() <= 5e-5


def test_metric_collection_shape_of_use(dev):
    """What compute_metric (core/utils.py:242-247) hands to the metrics: clamped sr and hr flattened to [(b t),c,h,w]."""
    from vsrlab_b200.losses import running_metrics_on_device
    g = torch.Generator().manual_seed(4)
    hr = torch.rand(2, 3, 3, 64, 64, generator=g).to(dev)
    sr = hr + 0.05 * torch.randn(2, 3, 3, 64, 64, generator=g).to(dev)
    m = running_metrics_on_device(sr, hr)
    x, y = sr.clamp(0, 1).flatten(0, 1), hr.flatten(0, 1)
    assert abs(m["PSNR"].item() - psnr_ref(x, y).mean().item()) <= 1e-4
    assert abs(m["SSIM"].item() - ssim_ref(x, y).mean().item()) <= 5e-5
