"""Runs the reference's own Real-BasicVSR modules (byte-compiled from /root/reference into oracle/_ref, see make_ref.py) on the CPU.
TEST / BASELINE INFRASTRUCTURE ONLY: used by `bench.py --impl reference` and its `cpu_baseline` leg, always in a process
of its own, because the reference calls itself `vsrlab` - the same package name as the drop-in.

    python -m oracle.ref_runner --blocks 5 --frames 3 --threads 16 [--seed 0] [--dump out.npz]

prints one JSON line {"frames": F, "seconds": S, "frames_per_s": V, "kind": "reference"}.
"""
from __future__ import annotations

import argparse
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import json
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "_ref" / "src"


class _BytecodeFinder(importlib.abc.MetaPathFinder):
    """Resolves `vsrlab[.sub.module]` to oracle/_ref/src/sub/module.bc (or .../__init__.bc for packages): the files
    make_ref.py byte-compiled from the reference, loaded with the stock sourceless loader."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "vsrlab" and not fullname.startswith("vsrlab."):
            return None
        rel = Path(*fullname.split(".")[1:])
        pkg, mod = SRC / rel / "__init__.bc", (SRC / rel).with_suffix(".bc")
        if pkg.exists():
            loader = importlib.machinery.SourcelessFileLoader(fullname, str(pkg))
            return importlib.util.spec_from_file_location(fullname, str(pkg), loader=loader, submodule_search_locations=[str(pkg.parent)])
        if mod.exists():
            loader = importlib.machinery.SourcelessFileLoader(fullname, str(mod))
            return importlib.util.spec_from_file_location(fullname, str(mod), loader=loader)
        return None


def load_reference_package():
    """Bind the package name `vsrlab` to oracle/_ref/src (recipe of SURVEY.md §8c)."""
    if any(k == "vsrlab" or k.startswith("vsrlab.") for k in sys.modules):
        raise RuntimeError("the drop-in `vsrlab` package is already imported in this process")
    sys.meta_path.insert(0, _BytecodeFinder())
    importlib.import_module("vsrlab")


def build_reference_model(blocks: int, seed: int = 0):
    import torch
    from vsrlab.vsr.models.RealBasicVSR.realbasicvsr import RealBasicVSR
    torch.manual_seed(seed)
    return RealBasicVSR(cleaning_blocks=blocks, mid_channels=64, upscale=4, res_blocks=blocks, pretrained_flow=False,
                        train_flow=False).eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=5)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--clips", type=int, default=1)
    ap.add_argument("--height", type=int, default=180)
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--dump", default="")
    a = ap.parse_args()
    import torch
    if a.threads > 0:
        torch.set_num_threads(a.threads)
    load_reference_package()
    net = build_reference_model(a.blocks, a.seed)
    x = torch.rand(a.clips, a.frames, 3, a.height, a.width, generator=torch.Generator().manual_seed(a.seed))
    with torch.no_grad():
        for _ in range(a.warmup):
            net(x[:, :2].clone())
        t0 = time.perf_counter()
        for _ in range(a.steps):
            sr, lq = net(x.clone())
        dt = time.perf_counter() - t0
    if a.dump:
        import numpy as np
        with torch.no_grad():                                  # the flows the forward used (basicvsr.py:43 on the cleaned clip)
            ff, fb = net.basicvsr.compute_flow(lq)
        np.savez(a.dump, sr=sr.numpy(), lq=lq.numpy(), flow_forward=ff.numpy(), flow_backward=fb.numpy())
    n = a.steps * a.clips * a.frames
    print(json.dumps({"frames": n, "seconds": dt, "frames_per_s": n / dt, "kind": "reference", "threads": torch.get_num_threads()}))


if __name__ == "__main__":
    main()
