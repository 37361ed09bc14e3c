"""PixelShufflePack with the reference's parameter layout
(reference src/core/modules/upsampling.py:4-12).  The shuffle is folded into the
conv kernel's store epilogue, so no shuffled copy is ever materialised."""
from torch import nn

from vsrlab_b200 import functional as VF


class PixelShufflePack(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, upscale_factor: int):
        super().__init__()
        r = self.upscale_factor = upscale_factor
        self.upconv = nn.Conv2d(in_ch, out_ch * r * r, kernel_size=3, stride=1, padding=1)
        self.pixel_shuffle = nn.PixelShuffle(r)     # kept for the parameter-free attribute; forward never calls it

    def forward(self, x):
        return VF.conv2d(x, self.upconv, act="none", pixel_shuffle=self.upscale_factor)
