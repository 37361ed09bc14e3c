import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    return load


def build_state_dict(kind):
    """Re-create the weights of a golden case with the drop-in package's constructors
    (seeded ctor == reference init, verified by checksum in test_mirror.py)."""
    import torch
    from vsrlab.vsr.models.RealBasicVSR.modules.spynet import Spynet
    from vsrlab.vsr.models.RealBasicVSR.realbasicvsr import RealBasicVSR

    def amplify(sd, prefix, gain):
        for lvl in range(6):
            for nm in ("weight", "bias"):
                k = f"{prefix}basic_module.{lvl}.basic_module.4.conv.0.{nm}"
                sd[k] = sd[k] * gain
        return sd

    if kind == "spynet":
        torch.manual_seed(21)
        return Spynet()
    if kind == "spynet_amp":
        torch.manual_seed(21)
        m = Spynet()
        m.load_state_dict(amplify({k: v.clone() for k, v in m.state_dict().items()}, "", 40.0))
        return m
    if kind == "cfg1":
        torch.manual_seed(0)
        return RealBasicVSR(cleaning_blocks=5, mid_channels=64, upscale=4, res_blocks=5, pretrained_flow=False, train_flow=True)
    if kind == "ragged":
        torch.manual_seed(5)
        m = RealBasicVSR(cleaning_blocks=2, mid_channels=64, upscale=4, res_blocks=2, pretrained_flow=False, train_flow=False)
        m.load_state_dict(amplify({k: v.clone() for k, v in m.state_dict().items()}, "basicvsr.spynet.", 40.0))
        return m
    raise KeyError(kind)
