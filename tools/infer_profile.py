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
    # idle time between consecutive kernels of the replay, grouped by (previous kernel -> next kernel)
    ks = sorted((e for e in prof.events() if str(e.device_type).endswith("CUDA") and e.time_range.end > e.time_range.start),
                key=lambda e: e.time_range.start)
    gaps = {}
    tot = 0.0
    for a_, b_ in zip(ks[:-1], ks[1:]):
        g = b_.time_range.start - a_.time_range.end
        if g <= 0:
            continue
        key = (a_.name[:44], b_.name[:44])
        d = gaps.setdefault(key, [0.0, 0])
        d[0] += g
        d[1] += 1
        tot += g
    print(f"idle between kernels: {tot / 1e3:.2f} ms over {len(ks)} kernels")
    for (pa, pb), (g, n) in sorted(gaps.items(), key=lambda kv: -kv[1][0])[:12]:
        print(f"{g / 1e3:7.3f} ms {n:4d}x {g / n:6.2f} us  {pa} -> {pb}")


if __name__ == "__main__" and "--replay-only" not in sys.argv:
    main()


def replay_only():
    """Time the captured graph alone (no copies in or out) against the public call."""
    from vsrlab_b200 import functional as VF
    dev = torch.device("cuda:0")
    VF.set_precision("bf16")
    model = bench.build_model(5, dev)
    lr = torch.rand(2, 30, 3, 180, 320, device=dev)
    with torch.no_grad():
        for _ in range(3):
            model(lr.clone())
    graph = next(iter(VF._graphs.values()))[1]
    for name, fn in (("graph.replay() only", graph.replay), ("model(lr)", lambda: model(lr))):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad():
            for _ in range(10):
                fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 10:.2f} ms per step")


if __name__ == "__main__" and "--replay-only" in sys.argv:
    replay_only()
