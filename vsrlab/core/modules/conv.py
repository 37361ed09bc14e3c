"""Hot-path building blocks with the reference's names and parameter layout
(reference src/core/modules/conv.py:15-22 ConvReLU, :82-92 ResidualConv,
:94-103 ResidualBlock).  Parameters stay fp32 OIHW ``nn.Conv2d`` tensors so
optimizers, DDP and checkpoints see exactly what they saw before; ``forward``
hands raw device pointers to the sm_100a kernels through the C-ABI.

Only the attribute names, the order in which the convolutions are created (it fixes the order of the RNG draws of a
seeded constructor) and the call signatures are the reference's; everything a forward does happens in vsrlab_b200."""
from torch import nn
from torch.nn.utils import spectral_norm

from vsrlab_b200 import functional as VF


def _conv3x3(c_in: int, c_out: int) -> nn.Conv2d:
    return nn.Conv2d(c_in, c_out, kernel_size=3, stride=1, padding=1)


class ConvReLU(nn.Module):
    """conv(k, stride 1) + ReLU.  Reference conv.py:15-22."""

    def __init__(self, in_ch, out_ch, *conv_args, **conv_kwargs):
        super().__init__()
        layers = [nn.Conv2d(in_ch, out_ch, *conv_args, **conv_kwargs), nn.ReLU()]
        self.conv = nn.Sequential(*layers)          # state_dict keys conv.0.{weight,bias}

    def forward(self, x):
        return VF.conv2d(x, self.conv[0], act="relu")


class ResidualConv(nn.Module):
    """x + conv2(relu(conv1(x))).  Reference conv.py:82-92."""

    def __init__(self, filters: int = 64):
        super().__init__()
        for name in ("conv1", "conv2"):             # created, hence initialised, in this order
            setattr(self, name, _conv3x3(filters, filters))
        self.relu = nn.ReLU()

    def forward(self, x):
        return VF.residual_stack(x, None, [self])


class ResidualBlock(nn.Module):
    """conv3x3 + LeakyReLU(0.1) stem, then `blocks` ResidualConvs.  Reference conv.py:94-103."""

    def __init__(self, in_ch: int, out_ch: int = 64, blocks: int = 30):
        super().__init__()
        stem = _conv3x3(in_ch, out_ch)
        self.conv = nn.Sequential(stem, nn.LeakyReLU(negative_slope=0.1))
        self.res_block = nn.Sequential(*(ResidualConv(out_ch) for _ in range(blocks)))

    def forward(self, x):
        return VF.residual_stack(x, self.conv[0], list(self.res_block))


class SpectralConv(nn.Module):
    """Bias-free conv under `torch.nn.utils.spectral_norm` (reference conv.py:6-13; the building block of the GAN
    discriminator, unet-discriminator.py:8-15).  Same parametrisation as the reference, hence the same state_dict keys
    (`conv.weight_orig`, `conv.weight_u`, `conv.weight_v`); 3x3 stride 1 and 4x4 stride 2 run on the sm_100a kernels."""

    def __init__(self, in_ch, out_ch, ks=3, stride=1, pad=1):
        super().__init__()
        self.conv = spectral_norm(nn.Conv2d(in_ch, out_ch, ks, stride, pad, bias=False))

    def forward(self, x):
        return VF.spectral_conv2d(x, self)
