"""Micro-benchmark of the tensor-core conv kernel: shapes x VSRB_TC_DEBUG modes, CUDA-event timed.

    python tools/conv_bench.py [--modes 0,1,2,3,4] [--cases hot]
"""
import argparse
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from vsrlab_b200 import ops  # noqa: E402
from vsrlab_b200._lib import ACT_NONE, ACT_RELU, BF16, EPI_SR  # noqa: E402

dev = torch.device("cuda:0")

CASES = {
    # name: (cin, cout, k, n, h, w, residual, pixshuf)
    "c64_3x3_6f": (64, 64, 3, 6, 180, 320, False, 0),
    "c64_3x3_6f_res": (64, 64, 3, 6, 180, 320, True, 0),
    "c64_3x3_2f": (64, 64, 3, 2, 180, 320, False, 0),
    "c64_3x3_30f": (64, 64, 3, 30, 180, 320, False, 0),
    "up_64_256_lr": (64, 256, 3, 2, 180, 320, False, 2),
    "up_64_256_2x": (64, 256, 3, 2, 360, 640, False, 2),
    "hr_64_64": (64, 64, 3, 2, 720, 1280, False, 0),
    "hr_64_3": (64, 3, 3, 2, 720, 1280, False, 0),
    "c128_3x3": (64, 128, 3, 6, 180, 320, False, 0),
    "sp_7x7_32_64": (32, 64, 7, 58, 96, 160, False, 0),
    "sp_7x7_64_32": (64, 32, 7, 58, 96, 160, False, 0),
    "sp_7x7_8_32": (8, 32, 7, 58, 192, 320, False, 0),
    "pt_1x1_128_64": (128, 64, 1, 6, 180, 320, False, 0),
    "stem_3_64": (3, 64, 3, 6, 180, 320, False, 0),
    "stem_3_64_30f": (3, 64, 3, 30, 180, 320, False, 0),
    "sr_64_3": (64, 3, 3, 8, 720, 1280, False, -1),     # conv_last: EPI_SR (fp32 NCHW out + bilinear x4 skip)
}


def run_case(name, reps=20):
    cin, cout, k, n, h, w, residual, ps = CASES[name]
    sr = ps < 0
    ps = max(ps, 0)
    cv = torch.nn.Conv2d(cin, cout, k, 1, k // 2).to(dev)
    pc = ops.PackedConv([cv], [(0, cin)], BF16, ps)
    ca = (cin + 15) // 16 * 16
    x = torch.randn(n, h, w, ca, device=dev).to(torch.bfloat16)
    r = ps or 1
    oc = max(16, (cout // (r * r) + 15) // 16 * 16) if ps else pc.cout_pad
    out = torch.empty(n, h * r, w * r, oc, dtype=torch.bfloat16, device=dev)
    res = torch.randn(n, h, w, oc, device=dev).to(torch.bfloat16) if residual else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    if sr:
        sr_out = torch.empty(n, cout, h, w, device=dev)
        lr = torch.randn(n, cout, h // 4, w // 4, device=dev)

    def call():
        if sr:
            ops.conv2d_fwd(pc, [x], [ca], n, h, w, act=ACT_NONE, epilogue=EPI_SR, f32_io=sr_out, f32_in=lr, aux_hw=(h // 4, w // 4))
            return
        ops.conv2d_fwd(pc, [x], [ca], n, h, w, act=ACT_RELU if not residual else ACT_NONE, out=out, out_c=oc, residual=res,
                       res_c=oc if residual else 0)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    flops = 2.0 * n * h * w * cout * cin * k * k
    return ms, flops / ms / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="0")
    ap.add_argument("--cases", default="all")
    ap.add_argument("--ring-modes", default="", help="VSRB_RING_DEBUG values to sweep instead of VSRB_TC_DEBUG (conv_ring.cu)")
    a = ap.parse_args()
    var = "VSRB_RING_DEBUG" if a.ring_modes else "VSRB_TC_DEBUG"
    if a.ring_modes:
        a.modes = a.ring_modes
    names = list(CASES) if a.cases == "all" else a.cases.split(",")
    print(f"{'case':18s} " + " ".join(f"{'mode' + m:>22s}" for m in a.modes.split(",")))
    for nm in names:
        row = []
        for m in a.modes.split(","):
            os.environ[var] = m
            ms, tf = run_case(nm)
            row.append(f"{ms*1e3:8.1f}us {tf:7.1f}TF/s")
        print(f"{nm:18s} " + " ".join(f"{r:>22s}" for r in row), flush=True)
    os.environ[var] = "0"
    assert ops.debug_status() == 0


if __name__ == "__main__":
    main()
