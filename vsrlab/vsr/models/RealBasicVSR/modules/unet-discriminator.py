"""U-Net discriminator of the Real-BasicVSR GAN recipe (reference
src/vsr/models/RealBasicVSR/modules/unet-discriminator.py:4-31; selected by conf/train/gan.yaml:17-20 through
`hydra.utils.instantiate`, which is why this file keeps the reference's hyphenated name).

Same attribute tree as the reference - `conv_0 ... conv_9`, spectral-norm parametrisation on conv_1 ... conv_8 - so
state_dict keys, seeded initialisation and optimizer parameter order are identical; the forward (three 4x4 stride-2
encoders, bilinear x2 decoders with skip additions, LeakyReLU 0.2) runs on the sm_100a kernels in vsrlab_b200."""
from torch import nn

from vsrlab.core.modules.conv import SpectralConv
from vsrlab_b200 import functional as VF


class UNetDiscriminator(nn.Module):
    def __init__(self, in_ch=3, mid_ch=64):
        super().__init__()
        widths = [mid_ch * 2, mid_ch * 4, mid_ch * 8]
        self.conv_0 = nn.Conv2d(in_ch, mid_ch, kernel_size=3, stride=1, padding=1)
        prev = mid_ch
        for i, c in enumerate(widths, start=1):                    # encoder: 4x4, stride 2
            setattr(self, f"conv_{i}", SpectralConv(prev, c, 4, 2, 1))
            prev = c
        for i, c in enumerate([mid_ch * 4, mid_ch * 2, mid_ch, mid_ch, mid_ch], start=4):   # decoder + head: 3x3, stride 1
            setattr(self, f"conv_{i}", SpectralConv(prev, c, 3, 1, 1))
            prev = c
        self.conv_9 = nn.Conv2d(mid_ch, 1, kernel_size=3, stride=1, padding=1)
        self.upsample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2)

    def forward(self, img):
        """img [B,3,H,W] (H, W multiples of 8) -> realness logits [B,1,H,W]."""
        return VF.unet_discriminator_forward(self, img)
