"""Launch the fused warp stem a few times (for `ncu --import-source on`): python tools/fused_stem_once.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vsrlab_b200 import ops  # noqa: E402
from vsrlab_b200._lib import ACT_LRELU, BF16  # noqa: E402

dev = torch.device("cuda:0")
n, h, w = 2, 180, 320
B = 2 * n
stem = ops.PackedConv([torch.nn.Conv2d(67, 64, 3, 1, 1).to(dev) for _ in range(2)], [(3, 64), (0, 3)], BF16)
feat = torch.randn(B, h, w, 64, device=dev).to(torch.bfloat16)
flow = (torch.rand(B, h, w, 2, device=dev) - 0.5) * 3.0
lr = torch.rand(B, h, w, 16, device=dev).to(torch.bfloat16)
patches = torch.randn(B, h, w, 32, device=dev).to(torch.bfloat16)
x1 = torch.empty_like(feat)
for _ in range(6):
    ops.conv2d_fwd(stem, [feat, lr], [64, 16], B, h, w, act=ACT_LRELU, out=x1, out_c=64, patch=patches, warp_flow=flow)
torch.cuda.synchronize()
assert ops.debug_status() == 0
print("ok")
