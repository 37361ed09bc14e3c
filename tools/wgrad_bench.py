"""Micro-benchmark of the weight-gradient kernels: python tools/wgrad_bench.py [--imgs 120 --hw 64]"""
import argparse
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from vsrlab_b200 import autograd as AG, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--imgs", type=int, default=120)
    ap.add_argument("--hw", type=int, default=64)
    ap.add_argument("--cin", type=int, default=64)
    ap.add_argument("--cout", type=int, default=64)
    ap.add_argument("--k", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cv = torch.nn.Conv2d(a.cin, a.cout, a.k, 1, a.k // 2)
    g = AG._wgrad_geom(cv, ((0, a.cin),))
    x = AG.to_cl16(torch.randn(a.imgs, a.cin, a.hw, a.hw).to(dev))
    dz = AG.to_cl16(torch.randn(a.imgs, a.cout, a.hw, a.hw).to(dev))
    dw = torch.zeros(a.cout, a.cin, a.k, a.k, device=dev)
    db = torch.zeros(a.cout, device=dev)

    def call():
        ops.conv2d_wgrad(g, [x], [x.shape[1]], dz, dz.shape[1], a.imgs, a.hw, a.hw, a.cin, dw, db)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    fl = 2.0 * a.imgs * a.hw * a.hw * a.cin * a.cout * a.k * a.k
    print(f"wgrad {a.k}x{a.k} {a.cin}->{a.cout} {a.imgs}x{a.hw}x{a.hw}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TF/s  "
          f"(VSRB_WG_DEBUG={os.environ.get('VSRB_WG_DEBUG', '0')} VSRB_WG_CTAS={os.environ.get('VSRB_WG_CTAS', '-')})")


if __name__ == "__main__":
    main()
