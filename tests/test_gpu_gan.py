"""GAN side network: the drop-in UNetDiscriminator / SpectralConv on the sm_100a conv kernels against
the golden outputs of the reference itself (tests/golden/gan.npz, made by make_golden.py) and against torch's own
conv2d on the same modules.  bf16 activations, so logits and gradients agree to bf16 noise (tolerances below)."""
import importlib

import numpy as np
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


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def make_D(dev, mid=16, seed=31):
    torch.manual_seed(seed)
    return importlib.import_module("vsrlab.vsr.models.RealBasicVSR.modules.unet-discriminator").UNetDiscriminator(3, mid).to(dev)


def torch_forward(D, img):
    """the reference forward (unet-discriminator.py:19-31) through torch's own conv2d on the same modules"""
    def sc(m, x):
        return m.conv(x)
    lr = lambda t: F.leaky_relu(t, 0.2)
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    f0 = lr(D.conv_0(img))
    f1 = lr(sc(D.conv_1, f0)); f2 = lr(sc(D.conv_2, f1)); f3 = up(lr(sc(D.conv_3, f2)))
    f4 = up(lr(sc(D.conv_4, f3)) + f2)
    f5 = up(lr(sc(D.conv_5, f4)) + f1)
    f6 = lr(sc(D.conv_6, f5)) + f0
    out = lr(sc(D.conv_8, lr(sc(D.conv_7, f6))))
    return D.conv_9(out)


@pytest.mark.parametrize("geom", [(8, 16, 3, 1, 1), (8, 16, 4, 2, 1), (64, 128, 4, 2, 1), (256, 128, 3, 1, 1)], ids=str)
def test_spectral_conv_matches_torch(dev, geom):
    from vsrlab.core.modules.conv import SpectralConv
    cin, cout, k, s, p = geom
    torch.manual_seed(3)
    m = SpectralConv(cin, cout, k, s, p).to(dev).eval()
    x = torch.randn(2, cin, 24, 40, device=dev).to(torch.bfloat16).float()
    with torch.no_grad():
        y = m(x)
        w = m.conv.weight.to(torch.bfloat16).float()         # eval mode: the hook leaves u, v alone
        ref = F.conv2d(x, w, None, s, p)
    assert y.shape == ref.shape
    assert rel(y, ref) < 6e-3                                # bf16 output rounding only
    # gradients reach weight_orig through the normalisation, and the input
    m.train()
    u0 = m.conv.weight_u.clone()
    xg = x.clone().requires_grad_(True)
    cot = torch.randn_like(ref)
    (m(xg) * cot).sum().backward()
    g_w, g_x = m.conv.weight_orig.grad.clone(), xg.grad.clone()
    assert not torch.equal(u0, m.conv.weight_u)               # one power iteration per training forward, like torch
    m.conv.weight_orig.grad = None
    m.conv.weight_u.copy_(u0)
    xr = x.clone().requires_grad_(True)
    (m.conv(xr) * cot).sum().backward()
    assert rel(g_x, xr.grad) < 2e-2 and rel(g_w, m.conv.weight_orig.grad) < 2e-2


def test_discriminator_matches_reference_golden(dev, golden):
    g = golden("gan")
    D = make_D(dev)
    img = torch.from_numpy(g["img"]).to(dev)
    D.eval()
    with torch.no_grad():
        y = D(img)
    assert y.shape == (2, 1, 32, 40) and y.dtype == torch.float32
    assert rel(y, g["logits_eval"]) < 2e-2, rel(y, g["logits_eval"])
    D.train()
    x = img.clone().requires_grad_(True)
    y = D(x)
    assert rel(y, g["logits_train"]) < 2e-2
    (y * torch.from_numpy(g["cot"]).to(dev)).sum().backward()
    np.testing.assert_allclose(D.conv_2.conv.weight_u.cpu().numpy(), g["u_after.conv_2"], atol=1e-5)
    # Gradients: bf16 noise through 10 layers each way, with heavy cancellation in the spectral-norm weight gradients.
    # The yardstick is torch's own bf16 autocast run of the reference forward on the same modules: this path must be
    # at least as close to the reference's fp32 gradients (measured: 6.6 % vs 7.3 % on the image, 12-14 % vs 15 %).
    D2 = make_D(dev).train()
    x2 = img.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y2 = torch_forward(D2, x2)
    (y2.float() * torch.from_numpy(g["cot"]).to(dev)).sum().backward()
    assert rel(x.grad, g["grad_img"]) < max(1.25 * rel(x2.grad, g["grad_img"]), 1e-2)
    assert F.cosine_similarity(x.grad.flatten().cpu(), torch.from_numpy(g["grad_img"]).flatten(), dim=0).item() > 0.995
    params, params2 = dict(D.named_parameters()), dict(D2.named_parameters())
    for k in ("conv_0.weight", "conv_0.bias", "conv_2.conv.weight_orig", "conv_5.conv.weight_orig", "conv_9.weight"):
        ours, theirs = rel(params[k].grad, g["grad." + k]), rel(params2[k].grad, g["grad." + k])
        assert ours < max(1.25 * theirs, 1e-2) and ours < 0.2, (k, ours, theirs)


def test_discriminator_full_width_and_weight_updates(dev):
    """mid_ch 64 at the GAN recipe's 256x256 crops: against torch on the GPU; packed weights follow optimizer steps."""
    D = make_D(dev, 64, seed=5)
    img = torch.rand(2, 3, 256, 256, device=dev)
    D.eval()
    with torch.no_grad():
        y, ref = D(img), torch_forward(D, img)
    assert rel(y, ref) < 2e-2, rel(y, ref)
    opt = torch.optim.Adam(D.parameters(), lr=1e-3)
    D.train()
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        F.binary_cross_entropy_with_logits(D(img), torch.ones_like(y)).backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in D.parameters())
        opt.step()
    D.eval()
    with torch.no_grad():
        y2, ref2 = D(img), torch_forward(D, img)
    assert rel(y2, ref2) < 2e-2 and rel(y2, y) > 1e-3          # follows the updated weights, not a stale packed image
    # the GAN recipe calls it under fp16 autocast with a GradScaler (train_gan.py:38-44): finite logits and gradients
    D.train()
    scaler = torch.amp.GradScaler("cuda")
    x = img.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        logits = D(x)
    assert logits.dtype == torch.float32 and torch.isfinite(logits).all()
    scaler.scale(F.binary_cross_entropy_with_logits(logits, torch.zeros_like(logits)) * 2e-5).backward()
    assert torch.isfinite(x.grad).all() and x.grad.abs().max() > 0


def reference_perceptual(vgg_layers, layer_weights, weight, yhat, y):
    """core/losses.py:39-64 verbatim in behaviour: tapped tensors are stored, then the next in-place ReLU runs on them"""
    def feats(x):
        out = {}
        for name, m in vgg_layers.named_children():
            x = m(x)
            if name in layer_weights:
                out[name] = x
        return out
    fx, fy = feats(yhat), feats(y.detach())
    return sum(F.l1_loss(fx[k], fy[k]) * w for k, w in layer_weights.items()) * weight


def test_perceptual_loss_on_the_conv_kernels(dev):
    """VGG19 features[:35] (random init: the pretrained blob needs a download) against torch on the same stack."""
    from torchvision import models
    from vsrlab_b200.losses import LAYER_WEIGHTS, AdversarialLoss, PerceptualLoss
    torch.manual_seed(9)
    vgg = models.vgg19(weights=None).features[:35].to(dev).eval()
    pl = PerceptualLoss(weight=1e-2, vgg_layers=vgg)
    assert not any(p.requires_grad for p in vgg.parameters())
    g = torch.Generator().manual_seed(10)
    hr = torch.rand(1, 2, 3, 64, 96, generator=g).to(dev)
    sr = (hr + 0.1 * torch.randn(hr.shape, generator=g).to(dev)).clamp(0, 1)
    a = sr.clone().requires_grad_(True)
    loss = pl(a, hr)
    loss.backward()
    b = sr.clone().requires_grad_(True)
    ref = reference_perceptual(vgg, LAYER_WEIGHTS, 1e-2, b.reshape(-1, 3, 64, 96), hr.reshape(-1, 3, 64, 96))
    ref.backward()
    c = sr.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        r16 = reference_perceptual(vgg, LAYER_WEIGHTS, 1e-2, c.reshape(-1, 3, 64, 96), hr.reshape(-1, 3, 64, 96))
    r16.float().backward()
    assert abs(loss.item() - ref.item()) < 2e-2 * abs(ref.item()), (loss.item(), ref.item())
    ours, theirs = rel(a.grad, b.grad), rel(c.grad, b.grad)
    # the L1 gradient is sign(fx - fy): bf16 rounding flips signs of near-equal features, for torch's bf16 run just as much
    # (measured 0.317 vs 0.315), so the yardstick is that run; the direction still agrees with the fp32 gradient
    assert ours < max(1.25 * theirs, 2e-2), (ours, theirs)
    assert F.cosine_similarity(a.grad.flatten(), b.grad.flatten(), dim=0).item() > 0.9
    adv = AdversarialLoss()
    x = torch.randn(4, 1, 8, 8, device=dev)
    assert torch.allclose(adv(x, 1, False), F.binary_cross_entropy_with_logits(x, torch.ones_like(x)) * 2e-5)
    assert torch.allclose(adv(x, 0, True), F.binary_cross_entropy_with_logits(x, torch.zeros_like(x)))


def test_two_forwards_before_backward_keep_their_own_weights(dev):
    """train_gan.py:49-58: D(hr) and D(sr) are both called before the backward.  In training mode each call moves the
    power-iteration vectors, so the two graphs hold different normalised weights - like torch, each backward must use its own."""
    from vsrlab.core.modules.conv import SpectralConv
    torch.manual_seed(4)
    m = SpectralConv(64, 64, 3, 1, 1).to(dev).train()
    ref = SpectralConv(64, 64, 3, 1, 1).to(dev).train()
    ref.load_state_dict(m.state_dict())
    x1 = torch.randn(2, 64, 16, 24, device=dev).to(torch.bfloat16).float()
    x2 = torch.randn(2, 64, 16, 24, device=dev).to(torch.bfloat16).float()
    a1, a2 = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    (m(a1).sum() + 3.0 * m(a2).sum()).backward()
    b1, b2 = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    (ref.conv(b1).sum() + 3.0 * ref.conv(b2).sum()).backward()
    assert rel(a1.grad, b1.grad) < 1e-2 and rel(a2.grad, b2.grad) < 1e-2
    assert rel(m.conv.weight_orig.grad, ref.conv.weight_orig.grad) < 2e-2
    # (the first call's input gradient must come from the FIRST normalised weight: at a random initialisation one power
    # iteration changes the weight by far more than the tolerance above)
    import gc
    from vsrlab_b200 import functional as VF, autograd as A
    n0 = (len(VF._packed), len(A._packed_t))
    for _ in range(5):
        m(x1).sum().backward()
    gc.collect()
    assert len(VF._packed) <= n0[0] + 1 and len(A._packed_t) <= n0[1] + 1          # per-call holders do not pile up
