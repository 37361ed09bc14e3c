"""Real-BasicVSR = pre-cleaning module + BasicVSR
(reference src/vsr/models/RealBasicVSR/realbasicvsr.py:5-15 RealBasicVSR,
:17-30 IterativeRefinement)."""
import torch.nn as nn

from vsrlab.core.modules.conv import ResidualBlock
from vsrlab.vsr.models.RealBasicVSR.modules.basicvsr import BasicVSR
from vsrlab_b200 import functional as VF


class RealBasicVSR(nn.Module):
    def __init__(self, cleaning_blocks=20, *args, **kwargs):
        super().__init__()
        # `mid_channels` must arrive as a kwarg, as in the reference (realbasicvsr.py:8)
        self.cleaner = IterativeRefinement(kwargs["mid_channels"], cleaning_blocks)
        self.basicvsr = BasicVSR(*args, **kwargs)

    def forward(self, lr):
        """lr [n,t,3,h,w] -> (sr, lq).  As in the reference (realbasicvsr.py:26-29) the
        cleaner works in place: `lq` aliases, and overwrites, the caller's `lr`."""
        return VF.realbasicvsr_forward(self, lr)


class IterativeRefinement(nn.Module):
    def __init__(self, mid_ch, blocks, steps=3):
        super().__init__()
        self.steps = steps
        self.resblock = ResidualBlock(3, mid_ch, blocks)
        self.conv = nn.Conv2d(mid_ch, 3, 3, 1, 1, bias=True)

    def forward(self, x):
        """x [n,t,3,h,w], refined in place `steps` times (realbasicvsr.py:24-30)."""
        return VF.cleaner_forward(self, x)
