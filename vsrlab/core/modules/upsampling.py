"""PixelShufflePack with the reference's parameter layout
(reference src/core/modules/upsampling.py:4-12).  The shuffle is folded into the
conv kernel's store epilogue, so no shuffled copy is ever materialised."""
import torch.nn as nn

from vsrlab_b200 import functional as VF


class PixelShufflePack(nn.Module):
    def __init__(self, in_ch, out_ch, upscale_factor):
        super().__init__()
        self.upconv = nn.Conv2d(in_ch, out_ch * upscale_factor * upscale_factor, 3, 1, 1)
        self.pixel_shuffle = nn.PixelShuffle(upscale_factor)
        self.upscale_factor = upscale_factor

    def forward(self, x):
        return VF.conv2d(x, self.upconv, act="none", pixel_shuffle=self.upscale_factor)
