"""Per-launch cost of the recurrent propagation convs the way the model runs them: a chain of dependent 3x3 64->64 convs
(conv+ReLU, conv+residual, ...) on a few images, two weight groups, programmatic dependent launch, replayed from a CUDA
graph.  Compares conv_ring_kernel with the stacked-layout kernel.

    python tools/chain_bench.py [--imgs 2] [--h 180 --w 320] [--layers 110]
"""
import argparse
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from vsrlab_b200 import ops  # noqa: E402
from vsrlab_b200._lib import ACT_NONE, ACT_RELU, BF16  # noqa: E402

dev = torch.device("cuda:0")


def run(a, env):
    for k in ("VSRB_TC_NO_RING", "VSRB_RING_MIN_ROWS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = a.groups
    B = a.imgs * g
    convs = [[torch.nn.Conv2d(64, 64, 3, 1, 1).to(dev) for _ in range(g)] for _ in range(2)]
    pcs = [ops.PackedConv(c, [(0, 64)], BF16) for c in convs]
    bufs = [torch.randn(B, a.h, a.w, 64, device=dev).to(torch.bfloat16) for _ in range(3)]

    def chain():
        cur = 0
        for j in range(a.layers // 2):
            t, o = (cur + 1) % 3, (cur + 2) % 3
            ops.conv2d_fwd(pcs[0], [bufs[cur]], [64], B, a.h, a.w, act=ACT_RELU, out=bufs[t], out_c=64)
            ops.conv2d_fwd(pcs[1], [bufs[t]], [64], B, a.h, a.w, act=ACT_NONE, out=bufs[o], out_c=64, residual=bufs[cur], res_c=64)
            cur = o
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        chain()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        chain()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    n = a.layers // 2 * 2
    flops = 2.0 * B * a.h * a.w * 64 * 64 * 9 * n
    assert ops.debug_status() == 0
    return ms * 1e3 / n, flops / ms / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--imgs", type=int, default=2, help="images per weight group")
    ap.add_argument("--groups", type=int, default=2)
    ap.add_argument("--h", type=int, default=180)
    ap.add_argument("--w", type=int, default=320)
    ap.add_argument("--layers", type=int, default=110)
    a = ap.parse_args()
    for name, env in (("ring", {"VSRB_RING_MIN_ROWS": "0"}), ("classic", {"VSRB_TC_NO_RING": "1"})):
        us, tf = run(a, env)
        print(f"{name:8s} imgs/group {a.imgs} groups {a.groups} {a.h}x{a.w}: {us:7.2f} us per launch  {tf:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
