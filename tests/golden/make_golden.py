"""Generate the golden fixtures in this directory from the REFERENCE ITSELF.

Runs only where /root/reference exists (the build container).  It imports the reference's
own modules (package alias `vsrlab`, recipe from SURVEY.md §8c), runs them on seeded inputs
on the CPU in fp32 and stores inputs + outputs as .npz.  The fixtures pin the CPU oracle
(`oracle/vsr_oracle.py`) and, through it and directly, the CUDA path.

    python tests/golden/make_golden.py

Weights are never stored: every case builds them with `torch.manual_seed(seed)` + the module
constructor, and the fixture records a checksum of the resulting state_dict so that the
drop-in `vsrlab` package (same constructor order => same RNG draws) can be checked to
initialise identically without the reference being present.
"""
import importlib.util
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/src")


def load_reference():
    for k in [k for k in sys.modules if k == "vsrlab" or k.startswith("vsrlab.")]:
        del sys.modules[k]
    spec = importlib.util.spec_from_file_location("vsrlab", str(REF / "__init__.py"), submodule_search_locations=[str(REF)])
    m = importlib.util.module_from_spec(spec)
    sys.modules["vsrlab"] = m
    spec.loader.exec_module(m)


def sd_checksum(sd):
    """order-independent fingerprint: per-tensor fp64 sum and sum of squares, keys sorted"""
    keys = sorted(sd.keys())
    return np.array([[sd[k].double().sum().item(), (sd[k].double() ** 2).sum().item()] for k in keys]), np.array(keys)


def amplify_flow(sd, prefix, gain):
    """scale the last conv of every SPyNet level so random-init flows reach several pixels"""
    for lvl in range(6):
        for nm in ("weight", "bias"):
            k = f"{prefix}basic_module.{lvl}.basic_module.4.conv.0.{nm}"
            sd[k] = sd[k] * gain
    return sd


def gan_golden():
    """UNetDiscriminator: eval-mode logits, train-mode logits (one power iteration) and gradients."""
    import importlib
    UNetDiscriminator = importlib.import_module("vsrlab.vsr.models.RealBasicVSR.modules.unet-discriminator").UNetDiscriminator
    out = {}
    torch.manual_seed(31)
    D = UNetDiscriminator(3, 16)
    cs, ck = sd_checksum(D.state_dict())
    out["sd_checksum"], out["sd_keys"] = cs, ck
    img = torch.rand(2, 3, 32, 40, generator=torch.Generator().manual_seed(32))
    out["img"] = img.numpy()
    D.eval()
    with torch.no_grad():
        out["logits_eval"] = D(img).numpy()
    D.train()
    x = img.clone().requires_grad_(True)
    y = D(x)
    out["logits_train"] = y.detach().numpy()
    w = torch.rand(y.shape, generator=torch.Generator().manual_seed(33)) - 0.5
    (y * w).sum().backward()
    out["cot"] = w.numpy()
    out["grad_img"] = x.grad.numpy()
    for k in ("conv_0.weight", "conv_0.bias", "conv_2.conv.weight_orig", "conv_5.conv.weight_orig", "conv_9.weight"):
        out["grad." + k] = dict(D.named_parameters())[k].grad.numpy()
    out["u_after." + "conv_2"] = D.conv_2.conv.weight_u.numpy()
    np.savez_compressed(HERE / "gan.npz", **out)


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(8)
    load_reference()
    if "--only-gan" in sys.argv:
        gan_golden()
        return
    gan_golden()
    from vsrlab.core.modules.conv import ResidualBlock
    from vsrlab.core.modules.upsampling import PixelShufflePack
    from vsrlab.vsr.models.RealBasicVSR.modules.basicvsr import BasicVSR
    from vsrlab.vsr.models.RealBasicVSR.modules.spynet import Spynet, flow_warp
    from vsrlab.vsr.models.RealBasicVSR.realbasicvsr import IterativeRefinement, RealBasicVSR

    # ---- small ops ----------------------------------------------------------------------
    g = torch.Generator().manual_seed(1234)
    out = {}
    x = torch.rand(2, 8, 12, 20, generator=g)
    fl = (torch.rand(2, 12, 20, 2, generator=g) - 0.5) * 30.0
    fl[0, 0, 0] = torch.tensor([0.0, 0.0]); fl[0, 0, 1] = torch.tensor([1.0, 2.0]); fl[0, 0, 2] = torch.tensor([0.5, -0.5])
    fl[0, 1, 0] = torch.tensor([-1.0, 0.0]); fl[0, 1, 1] = torch.tensor([19.0, 11.0]); fl[0, 1, 2] = torch.tensor([100.0, -100.0])
    fl[1, 11, 19] = torch.tensor([0.0, 0.0]); fl[1, 11, 18] = torch.tensor([1.0, 0.0]); fl[1, 11, 17] = torch.tensor([2.5, 0.5])
    out["warp_x"], out["warp_flow"] = x.numpy(), fl.numpy()
    out["warp_zeros"] = flow_warp(x, fl).numpy()
    out["warp_border"] = flow_warp(x, fl, padding_mode="border").numpy()

    torch.manual_seed(11)
    rb = ResidualBlock(3, 16, 2).eval()
    xi = torch.rand(2, 3, 9, 14)
    with torch.no_grad():
        out["rb_x"], out["rb_y"] = xi.numpy(), rb(xi).numpy()
    for k, v in rb.state_dict().items():
        out["rb_sd." + k] = v.numpy()

    torch.manual_seed(12)
    ps = PixelShufflePack(16, 16, 2).eval()
    xi = torch.rand(1, 16, 6, 10)
    with torch.no_grad():
        out["ps_x"], out["ps_y"] = xi.numpy(), ps(xi).numpy()
    for k, v in ps.state_dict().items():
        out["ps_sd." + k] = v.numpy()

    torch.manual_seed(13)
    ir = IterativeRefinement(16, 1).eval()
    xi = torch.rand(1, 2, 3, 10, 12)
    out["ir_x"] = xi.numpy().copy()
    with torch.no_grad():
        y = ir(xi)
    out["ir_y"] = y.numpy()
    out["ir_aliases_input"] = np.array(y.data_ptr() == xi.data_ptr())
    for k, v in ir.state_dict().items():
        out["ir_sd." + k] = v.numpy()
    np.savez_compressed(HERE / "ops.npz", **out)

    # ---- SPyNet: exact-/32 and ragged sizes, default and amplified flows -----------------
    out = {}
    torch.manual_seed(21)
    sp = Spynet().eval()
    cs, ck = sd_checksum(sp.state_dict())
    out["sd_checksum"], out["sd_keys"] = cs, ck
    g = torch.Generator().manual_seed(22)
    for tag, shape in (("a", (2, 3, 64, 96)), ("b", (3, 3, 36, 52))):
        r = torch.rand(*shape, generator=g)
        s = torch.rand(*shape, generator=g)
        with torch.no_grad():
            out[f"{tag}_ref"], out[f"{tag}_supp"], out[f"{tag}_flow"] = r.numpy(), s.numpy(), sp(r, s).numpy()
    sp.load_state_dict(amplify_flow({k: v.clone() for k, v in sp.state_dict().items()}, "", 40.0))
    for tag, shape in (("c", (2, 3, 36, 52)),):
        r = torch.rand(*shape, generator=g)
        s = torch.rand(*shape, generator=g)
        with torch.no_grad():
            out[f"{tag}_ref"], out[f"{tag}_supp"], out[f"{tag}_flow"] = r.numpy(), s.numpy(), sp(r, s).numpy()
    np.savez_compressed(HERE / "spynet.npz", **out)

    # ---- cfg1: Real-BasicVSR x4, experiment=basic (5/5 blocks), 1x5x3x64x64 ------------
    out = {}
    torch.manual_seed(0)
    net = RealBasicVSR(cleaning_blocks=5, mid_channels=64, upscale=4, res_blocks=5, pretrained_flow=False, train_flow=True).eval()
    cs, ck = sd_checksum(net.state_dict())
    out["sd_checksum"], out["sd_keys"] = cs, ck
    x = torch.rand(1, 5, 3, 64, 64)
    out["lr"] = x.numpy().copy()
    with torch.no_grad():
        ff, fb = net.basicvsr.compute_flow(net.cleaner(x.clone()))
        sr, lq = net(x)
    out["sr"], out["lq"] = sr.numpy(), lq.numpy()
    out["flow_forward"], out["flow_backward"] = ff.numpy(), fb.numpy()
    out["lq_aliases_input"] = np.array(lq.data_ptr() == x.data_ptr())
    out["fingerprint"] = np.array([sr.mean().item(), sr.std().item(), sr.abs().max().item()])
    np.savez_compressed(HERE / "cfg1.npz", **out)

    # ---- ragged: 2 clips x 3 frames 36x52, 2/2 blocks, amplified flows, BasicVSR alone too
    out = {}
    torch.manual_seed(5)
    net = RealBasicVSR(cleaning_blocks=2, mid_channels=64, upscale=4, res_blocks=2, pretrained_flow=False, train_flow=False).eval()
    net.load_state_dict(amplify_flow({k: v.clone() for k, v in net.state_dict().items()}, "basicvsr.spynet.", 40.0))
    cs, ck = sd_checksum(net.state_dict())
    out["sd_checksum"], out["sd_keys"] = cs, ck
    x = torch.rand(2, 3, 3, 36, 52)
    out["lr"] = x.numpy().copy()
    with torch.no_grad():
        out["basicvsr_sr"] = net.basicvsr(x.clone()).numpy()
        ff, fb = net.basicvsr.compute_flow(x.clone())
        out["basicvsr_flow_forward"], out["basicvsr_flow_backward"] = ff.numpy(), fb.numpy()
        sr, lq = net(x)
    out["sr"], out["lq"] = sr.numpy(), lq.numpy()
    np.savez_compressed(HERE / "ragged.npz", **out)

    # ---- cfg2 (BASELINE.json configs[1]): SPyNet on a synthetic 7-frame 256x256 clip, both pair directions.
    # The clip is regenerated from its seed by the tests; the flows are stored on a stride-2 lattice (phase (0,1))
    # to keep the fixture small, plus whole-field statistics.
    out = {}
    torch.manual_seed(21)
    sp = Spynet().eval()
    clip = torch.rand(7, 3, 256, 256, generator=torch.Generator().manual_seed(2024))
    with torch.no_grad():
        fb = sp(clip[:-1], clip[1:])          # basicvsr.py:35 (backward flows)
        ff = sp(clip[1:], clip[:-1])          # basicvsr.py:36 (forward flows)
    for nm, f in (("backward", fb), ("forward", ff)):
        out[f"flow_{nm}_s2"] = f[:, :, 0::2, 1::2].contiguous().numpy()
        out[f"flow_{nm}_stats"] = np.array([f.double().mean().item(), f.double().std().item(), f.abs().max().item()])
    out["clip_checksum"] = np.array([clip.double().sum().item()])
    np.savez_compressed(HERE / "cfg2.npz", **out)
    for f in sorted(HERE.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
