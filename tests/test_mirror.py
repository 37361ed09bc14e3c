"""Host-side contract of the drop-in `vsrlab` package (no GPU needed)."""
import numpy as np
import pytest
import torch

from conftest import build_state_dict


@pytest.mark.parametrize("kind,fixture", [("cfg1", "cfg1"), ("ragged", "ragged"), ("spynet", "spynet")])
def test_seeded_ctor_reproduces_reference_init(golden, kind, fixture):
    """Same constructor order => same RNG draws => bit-identical parameters."""
    g = golden(fixture)
    sd = build_state_dict(kind).state_dict()
    keys = sorted(sd.keys())
    assert keys == [str(k) for k in g["sd_keys"]]
    cs = np.array([[sd[k].double().sum().item(), (sd[k].double() ** 2).sum().item()] for k in keys])
    np.testing.assert_allclose(cs, g["sd_checksum"], rtol=0, atol=1e-9)


def test_discriminator_ctor_reproduces_reference_init(golden):
    """UNetDiscriminator (gan.yaml:17-20): same keys (spectral-norm parametrisation included) and seeded values."""
    import importlib
    g = golden("gan")
    torch.manual_seed(31)
    D = importlib.import_module("vsrlab.vsr.models.RealBasicVSR.modules.unet-discriminator").UNetDiscriminator(3, 16)
    sd = D.state_dict()
    keys = sorted(sd.keys())
    assert keys == [str(k) for k in g["sd_keys"]]
    cs = np.array([[sd[k].double().sum().item(), (sd[k].double() ** 2).sum().item()] for k in keys])
    np.testing.assert_allclose(cs, g["sd_checksum"], rtol=0, atol=1e-9)
    assert [n for n, _ in D.named_parameters()][:3] == ["conv_0.weight", "conv_0.bias", "conv_1.conv.weight_orig"]
    with pytest.raises(Exception):                       # no CPU fallback here either
        D(torch.rand(1, 3, 16, 16))


def test_state_dict_layout_and_ctor_contract():
    from vsrlab.vsr.models.RealBasicVSR.realbasicvsr import RealBasicVSR
    m = RealBasicVSR(cleaning_blocks=5, mid_channels=64, upscale=4, res_blocks=5, pretrained_flow=False, train_flow=True)
    sd = m.state_dict()
    assert len(sd) == 140 and sum(p.numel() for p in m.parameters()) == 2971314
    for k in ("cleaner.resblock.conv.0.weight", "cleaner.resblock.res_block.4.conv2.bias", "cleaner.conv.weight",
              "basicvsr.backward_resblocks.conv.0.weight", "basicvsr.forward_resblocks.res_block.0.conv1.weight",
              "basicvsr.point_conv.0.weight", "basicvsr.upsample.1.upconv.weight", "basicvsr.conv_last.2.bias",
              "basicvsr.spynet.basic_module.5.basic_module.4.conv.0.weight", "basicvsr.spynet.mean", "basicvsr.spynet.std"):
        assert k in sd
    assert sd["basicvsr.backward_resblocks.conv.0.weight"].shape == (64, 67, 3, 3)
    assert sd["basicvsr.point_conv.0.weight"].shape == (64, 128, 1, 1)
    assert sd["basicvsr.upsample.0.upconv.weight"].shape == (256, 64, 3, 3)
    with pytest.raises(KeyError):                       # mid_channels must be a kwarg (realbasicvsr.py:8)
        RealBasicVSR(5)
    frozen = RealBasicVSR(cleaning_blocks=1, mid_channels=64, res_blocks=1, train_flow=False)
    assert not any(p.requires_grad for p in frozen.basicvsr.spynet.parameters())     # basicvsr.py:25-28
    assert all(p.requires_grad for p in m.basicvsr.spynet.parameters())


def test_no_cpu_fallback():
    """The product path must fail loudly off-GPU instead of silently computing on the CPU."""
    from vsrlab.vsr.models.RealBasicVSR.realbasicvsr import RealBasicVSR
    from vsrlab.vsr.models.RealBasicVSR.modules.spynet import flow_warp
    from vsrlab_b200 import VsrbError
    m = RealBasicVSR(cleaning_blocks=1, mid_channels=64, res_blocks=1).eval()
    with torch.no_grad(), pytest.raises(VsrbError):
        m(torch.rand(1, 2, 3, 8, 8))
    with pytest.raises(VsrbError):
        flow_warp(torch.rand(1, 8, 4, 4), torch.zeros(1, 4, 4, 2))


def test_product_path_does_not_import_oracle():
    import pathlib
    root = pathlib.Path(__file__).resolve().parents[1]
    for pkg in ("vsrlab", "vsrlab_b200"):
        for f in (root / pkg).rglob("*.py"):
            src = f.read_text()
            assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"
