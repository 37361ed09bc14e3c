"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total time, share.

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rN_launch_summary.csv
"""
import csv
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        t = tot[r[ik][:70]]
        t[0] += 1
        t[1] += v * scale
    allms = sum(t[1] for t in tot.values())
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "total_ms", "share", "avg_us"])
    for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, n, f"{ms:.3f}", f"{ms / allms:.4f}", f"{ms * 1e3 / n:.1f}"])


if __name__ == "__main__":
    main()
