"""BasicVSR with the reference's parameter tree
(reference src/vsr/models/RealBasicVSR/modules/basicvsr.py:11-28 ctor, :30-37
compute_flow, :39-83 forward)."""
import logging

from torch import nn

from vsrlab.core.modules.conv import ResidualBlock
from vsrlab.core.modules.upsampling import PixelShufflePack
from vsrlab.vsr.models.RealBasicVSR.modules.spynet import Spynet
from vsrlab_b200 import functional as VF

log = logging.getLogger(__name__)


class BasicVSR(nn.Module):
    def __init__(self, mid_channels: int = 64, res_blocks: int = 30, upscale: int = 4, pretrained_flow: bool = False,
                 train_flow: bool = False):
        super().__init__()
        c = self.mid_channels = mid_channels
        # Sub-modules are created in the reference's order (a seeded constructor then draws the same initial weights) and
        # under the reference's attribute names (the state_dict keys); what they compute lives in vsrlab_b200.functional.
        for direction in ("backward", "forward"):                      # one trunk per propagation direction
            setattr(self, f"{direction}_resblocks", ResidualBlock(c + 3, c, res_blocks))
        self.point_conv = nn.Sequential(nn.Conv2d(2 * c, c, kernel_size=1), nn.LeakyReLU(negative_slope=0.1))
        self.upsample = nn.Sequential(*(PixelShufflePack(c, c, 2) for _ in range(upscale // 2)))
        self.conv_last = nn.Sequential(nn.Conv2d(c, 64, kernel_size=3, padding=1), nn.LeakyReLU(negative_slope=0.1),
                                       nn.Conv2d(64, 3, kernel_size=3, padding=1))
        self.upscale = nn.Upsample(scale_factor=upscale, mode="bilinear", align_corners=False)
        self.spynet = Spynet(pretrained_flow)
        if not train_flow:                                               # frozen flow network (basicvsr.py:25-28)
            log.info("optical-flow weights frozen (train_flow=False)")
            self.spynet.requires_grad_(False)

    def compute_flow(self, lrs):
        """(flow_forward, flow_backward), each [n*(t-1), 2, h, w] (basicvsr.py:30-37)."""
        return VF.basicvsr_flows(self, lrs)

    def forward(self, lrs):
        """lrs [n,t,3,h,w] -> sr [n,t,3,s*h,s*w] (basicvsr.py:39-83)."""
        return VF.basicvsr_forward(self, lrs)
