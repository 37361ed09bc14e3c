"""Timeline of one small tensor-core conv launch (VSRB_TC_DEBUG=64 -> vsrb_debug_trace): where do the ~10 us between
`tiles x steady-state tile time` and the measured launch time go?

    python tools/trace_small.py [--frames 4] [--groups 2]
"""
import argparse
import ctypes as C
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ["VSRB_TC_DEBUG"] = "64"
from vsrlab_b200 import ops  # noqa: E402
from vsrlab_b200 import _lib as L  # noqa: E402
from vsrlab_b200._lib import ACT_RELU, BF16  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--groups", type=int, default=2)
    ap.add_argument("--chain", type=int, default=6, help="back-to-back launches before the traced one")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n, h, w, g = a.frames, 180, 320, a.groups
    convs = [torch.nn.Conv2d(64, 64, 3, 1, 1).to(dev) for _ in range(g)]
    pc = ops.PackedConv(convs, [(0, 64)], BF16, 0)
    x = torch.randn(n, h, w, 64, device=dev).to(torch.bfloat16)
    y = torch.empty_like(x)

    def call(src, dst):
        ops.conv2d_fwd(pc, [src], [64], n, h, w, imgs_per_group=n // g, act=ACT_RELU, out=dst, out_c=64)
    for _ in range(3):
        call(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.chain):
        call(x, y) if i % 2 == 0 else call(y, x)
    e1.record()
    torch.cuda.synchronize()
    print(f"{a.chain} chained launches: {e0.elapsed_time(e1) * 1e3 / a.chain:.1f} us per launch")
    # the same chain as a CUDA graph (what the model's forward replays)
    gr = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr, stream=side):
            for i in range(40):
                call(x, y) if i % 2 == 0 else call(y, x)
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"graph of 40 chained launches: {e0.elapsed_time(e1) * 1e3 / 200:.1f} us per launch")
    nct = 148
    buf = (C.c_uint64 * (nct * 8))()
    L.check(L.load().vsrb_debug_trace(buf, nct), "trace")
    rows = [[buf[c * 8 + i] for i in range(8)] for c in range(nct)]
    t0 = min(r[0] for r in rows if r[0])
    names = ["entry", "prologue", "weights", "tile0 loaded", "acc0 ready", "last tile", "stores drained", "exit"]
    print("stamps in us relative to the first CTA's entry (cols: " + ", ".join(names) + ")")
    for c in (0, 1, 2, 73, 74, 146, 147):
        print(f"cta {c:3d}: " + " ".join(f"{(v - t0) / 1e3:7.2f}" if v else "      -" for v in rows[c]))
    import statistics as st
    for i, nm in enumerate(names):
        vals = [(r[i] - t0) / 1e3 for r in rows if r[i]]
        if vals:
            print(f"{nm:15s} min {min(vals):7.2f}  median {st.median(vals):7.2f}  max {max(vals):7.2f}  (n={len(vals)})")


if __name__ == "__main__":
    main()
