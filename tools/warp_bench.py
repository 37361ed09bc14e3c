"""flow_warp micro-benchmark on a working set larger than L2 (CUDA events, L2 flushed between reps)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from vsrlab_b200 import ops  # noqa: E402
from vsrlab_b200._lib import BF16, F32, PAD_ZEROS  # noqa: E402

dev = torch.device("cuda:0")


def run(n, h, w, c, dt, mag, reps=10):
    tdt = ops.TORCH_DT[dt]
    x = torch.randn(n, h, w, c, device=dev).to(tdt)
    fl = (torch.rand(n, h, w, 2, device=dev) - 0.5) * 2 * mag
    out = torch.empty_like(x)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(3):
        ops.flow_warp(x, fl, out, n, h, w, c, dt, PAD_ZEROS)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.flow_warp(x, fl, out, n, h, w, c, dt, PAD_ZEROS)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    byt = n * h * w * (2 * c * ops.ESIZE[dt] + 8)
    return ms, byt / ms / 1e6


if __name__ == "__main__":
    for n in (2, 8, 64, 256):
        for mag in (0.8, 8.0):
            ms, gbs = run(n, 180, 320, 64, BF16, mag)
            print(f"bf16 n={n:4d} |flow|<={mag:4.1f}px  {ms*1e3:8.1f} us  {gbs:8.1f} GB/s algorithmic", flush=True)
    ms, gbs = run(64, 180, 320, 64, F32, 0.8)
    print(f"f32  n=  64 |flow|<= 0.8px  {ms*1e3:8.1f} us  {gbs:8.1f} GB/s algorithmic")
