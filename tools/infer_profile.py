"""CUPTI (torch.profiler) table of every kernel of one replayed cfg3 inference step: the GPU time of each kernel class
inside the CUDA graph, without the launch gaps an eager per-launch timing adds.   python tools/infer_profile.py [--clips 2]"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=2)
    ap.add_argument("--blocks", type=int, default=5)
    a = ap.parse_args()
    from torch.profiler import ProfilerActivity, profile
    from vsrlab_b200 import functional as VF
    dev = torch.device("cuda:0")
    VF.set_precision("bf16")
    model = bench.build_model(a.blocks, dev)
    lr = torch.rand(a.clips, 30, 3, 180, 320, device=dev)
    with torch.no_grad():
        for _ in range(4):
            model(lr.clone())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            model(lr.clone())
        e1.record()
        torch.cuda.synchronize()
        print(f"step {e0.elapsed_time(e1) / 5:.2f} ms")
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            model(lr.clone())
            torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=80))


if __name__ == "__main__":
    main()
