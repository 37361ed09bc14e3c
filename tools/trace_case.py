"""Pipeline milestones (VSRB_TC_DEBUG=64 -> vsrb_debug_trace) of one conv_bench CASE, optionally with other debug bits
(1 / 2 / 4 / 8 = no loads / stores / MMAs / epilogue), plus the launch's grid and tile plan.

    python tools/trace_case.py --case hr_64_3 [--extra 13]
"""
import argparse
import ctypes as C
import os
import statistics as st
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="hr_64_3")
    ap.add_argument("--extra", type=int, default=0)
    a = ap.parse_args()
    os.environ["VSRB_TC_DEBUG"] = str(64 | 32 | a.extra)
    import torch
    import conv_bench as CB
    from vsrlab_b200 import _lib as L
    t = CB.run_case(a.case, reps=5)
    torch.cuda.synchronize()
    print(f"{a.case} debug {64 | a.extra}: {t}")
    nct = 512
    buf = (C.c_uint64 * (nct * 8))()
    L.check(L.load().vsrb_debug_trace(buf, nct), "trace")
    rows = [[buf[c * 8 + i] for i in range(8)] for c in range(148)]
    pr, mm, ep = ([buf[r * 8 + i] for i in range(3)] for r in (300, 301, 302))
    print(f"CTA 0 cycles: producer total {pr[0]} waiting for a free slot {pr[1]} | MMA warp total {mm[0]} waiting for an accumulator {mm[1]} "
          f"for operands {mm[2]} | epilogue warp 4 total {ep[0]} waiting for an accumulator {ep[1]}")
    t0 = min(r[0] for r in rows if r[0])
    names = ["entry", "prologue", "weights", "tile0 loaded", "acc0 ready", "last tile", "stores drained", "exit"]
    for i, nm in enumerate(names):
        vals = [(r[i] - t0) / 1e3 for r in rows if r[i]]
        if vals:
            print(f"{nm:15s} min {min(vals):7.2f}  median {st.median(vals):7.2f}  max {max(vals):7.2f}  (n={len(vals)})")


if __name__ == "__main__":
    main()
