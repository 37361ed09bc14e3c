"""BasicVSR with the reference's parameter tree
(reference src/vsr/models/RealBasicVSR/modules/basicvsr.py:11-28 ctor, :30-37
compute_flow, :39-83 forward)."""
import logging

import torch.nn as nn

from vsrlab.core.modules.conv import ResidualBlock
from vsrlab.core.modules.upsampling import PixelShufflePack
from vsrlab.vsr.models.RealBasicVSR.modules.spynet import Spynet
from vsrlab_b200 import functional as VF

pylogger = logging.getLogger(__name__)


class BasicVSR(nn.Module):
    def __init__(self, mid_channels=64, res_blocks=30, upscale=4,
                 pretrained_flow=False, train_flow=False):
        super().__init__()
        self.mid_channels = mid_channels
        # construction order == reference order, so a seeded ctor draws the same init
        self.backward_resblocks = ResidualBlock(mid_channels + 3, mid_channels, res_blocks)
        self.forward_resblocks = ResidualBlock(mid_channels + 3, mid_channels, res_blocks)
        self.point_conv = nn.Sequential(nn.Conv2d(mid_channels * 2, mid_channels, 1, 1), nn.LeakyReLU(0.1))
        self.upsample = nn.Sequential(*[PixelShufflePack(mid_channels, mid_channels, 2) for _ in range(upscale // 2)])
        self.conv_last = nn.Sequential(nn.Conv2d(mid_channels, 64, 3, 1, 1), nn.LeakyReLU(0.1),
                                       nn.Conv2d(64, 3, 3, 1, 1))
        self.upscale = nn.Upsample(scale_factor=upscale, mode='bilinear', align_corners=False)
        self.spynet = Spynet(pretrained_flow)

        if not train_flow:
            pylogger.info('Setting Optical Flow weights to no_grad')
            for param in self.spynet.parameters():
                param.requires_grad = False

    def compute_flow(self, lrs):
        """(flow_forward, flow_backward), each [n*(t-1), 2, h, w] (basicvsr.py:30-37)."""
        return VF.basicvsr_flows(self, lrs)

    def forward(self, lrs):
        """lrs [n,t,3,h,w] -> sr [n,t,3,s*h,s*w] (basicvsr.py:39-83)."""
        return VF.basicvsr_forward(self, lrs)
