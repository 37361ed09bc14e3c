"""Hot-path building blocks with the reference's names and parameter layout
(reference src/core/modules/conv.py:15-22 ConvReLU, :82-92 ResidualConv,
:94-103 ResidualBlock).  Parameters stay fp32 OIHW ``nn.Conv2d`` tensors so
optimizers, DDP and checkpoints see exactly what they saw before; ``forward``
hands raw device pointers to the sm_100a kernels through the C-ABI."""
import torch.nn as nn

from vsrlab_b200 import functional as VF


class ConvReLU(nn.Module):
    """conv(k, stride 1) + ReLU.  Reference conv.py:15-22."""

    def __init__(self, in_ch, out_ch, *args, **kwargs):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_ch, out_ch, *args, **kwargs), nn.ReLU())

    def forward(self, x):
        return VF.conv2d(x, self.conv[0], act="relu")


class ResidualConv(nn.Module):
    """x + conv2(relu(conv1(x))).  Reference conv.py:82-92."""

    def __init__(self, filters=64):
        super().__init__()
        self.conv1 = nn.Conv2d(filters, filters, 3, 1, 1)
        self.conv2 = nn.Conv2d(filters, filters, 3, 1, 1)
        self.relu = nn.ReLU()

    def forward(self, x):
        return VF.residual_stack(x, None, [self])


class ResidualBlock(nn.Module):
    """conv3x3 + LeakyReLU(0.1) stem, then `blocks` ResidualConvs.  Reference conv.py:94-103."""

    def __init__(self, in_ch, out_ch=64, blocks=30):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_ch, out_ch, 3, 1, 1), nn.LeakyReLU(0.1))
        self.res_block = nn.Sequential(*[ResidualConv(out_ch) for _ in range(blocks)])

    def forward(self, x):
        return VF.residual_stack(x, self.conv[0], list(self.res_block))
