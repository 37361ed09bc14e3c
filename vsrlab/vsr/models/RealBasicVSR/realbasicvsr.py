"""Real-BasicVSR = pre-cleaning module + BasicVSR
(reference src/vsr/models/RealBasicVSR/realbasicvsr.py:5-15 RealBasicVSR,
:17-30 IterativeRefinement)."""
from torch import nn

from vsrlab.core.modules.conv import ResidualBlock
from vsrlab.vsr.models.RealBasicVSR.modules.basicvsr import BasicVSR
from vsrlab_b200 import functional as VF


class RealBasicVSR(nn.Module):
    def __init__(self, cleaning_blocks: int = 20, *basicvsr_args, **basicvsr_kwargs):
        super().__init__()
        # `mid_channels` must arrive as a keyword, as in the reference (realbasicvsr.py:8): the cleaner is sized from it
        width = basicvsr_kwargs["mid_channels"]
        self.cleaner = IterativeRefinement(width, cleaning_blocks)
        self.basicvsr = BasicVSR(*basicvsr_args, **basicvsr_kwargs)

    def forward(self, lr):
        """lr [n,t,3,h,w] -> (sr, lq).  As in the reference (realbasicvsr.py:26-29) the
        cleaner works in place: `lq` aliases, and overwrites, the caller's `lr`."""
        return VF.realbasicvsr_forward(self, lr)


class IterativeRefinement(nn.Module):
    def __init__(self, mid_ch: int, blocks: int, steps: int = 3):
        super().__init__()
        self.steps = steps
        self.resblock = ResidualBlock(3, mid_ch, blocks)                               # image -> features
        self.conv = nn.Conv2d(mid_ch, 3, kernel_size=3, stride=1, padding=1, bias=True)   # features -> residue

    def forward(self, x):
        """x [n,t,3,h,w], refined in place `steps` times (realbasicvsr.py:24-30)."""
        return VF.cleaner_forward(self, x)
