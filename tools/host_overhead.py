"""Host-side cost of one library call (the training path is launch-bound on 64x64 patches):

    python tools/host_overhead.py

prints the time per call of ops.conv2d_fwd (ctypes struct fill + C call), of the bare C call, and of the autograd wrappers
(forward, forward + backward) on an 8x8 image where the GPU work is negligible."""
import ctypes as C
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from vsrlab_b200 import _lib as L  # noqa: E402
from vsrlab_b200 import autograd as AG, ops  # noqa: E402
from vsrlab_b200._lib import ACT_RELU, BF16  # noqa: E402


def timed(fn, n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6


def main():
    dev = torch.device("cuda:0")
    cv = torch.nn.Conv2d(64, 64, 3, 1, 1).to(dev)
    pc = ops.PackedConv([cv], [(0, 64)], BF16, 0)
    x = torch.randn(1, 8, 8, 64, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)

    def call():
        ops.conv2d_fwd(pc, [x], [64], 1, 8, 8, act=ACT_RELU, out=y, out_c=64)
    for _ in range(10):
        call()
    print(f"ops.conv2d_fwd:              {timed(call, 2000):6.1f} us per call (host)")
    a = L.ConvArgs()
    a.geom = pc.geom
    a.inp[0], a.in_c[0] = x.data_ptr(), 64
    a.batch, a.h, a.w, a.imgs_per_group = 1, 8, 8, 1
    a.packed, a.act, a.slope, a.out, a.out_c = pc.buf.data_ptr(), ACT_RELU, 0.1, y.data_ptr(), 64
    lib, st = L.load(), torch.cuda.current_stream().cuda_stream
    print(f"vsrb_conv2d_fwd (C call):    {timed(lambda: lib.vsrb_conv2d_fwd(C.byref(a), st), 2000):6.1f} us")
    xx = torch.randn(1, 64, 8, 8, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g = torch.ones(1, 64, 8, 8, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)

    def fwd():
        return AG.conv(cv, [xx], [(0, 64)], "relu")

    def fwd_bwd():
        fwd().backward(g)
    for _ in range(20):
        fwd_bwd()
    print(f"autograd conv forward:       {timed(fwd, 500):6.1f} us")
    print(f"autograd conv fwd + bwd:     {timed(fwd_bwd, 500):6.1f} us")


if __name__ == "__main__":
    main()
