"""BASELINE cfg2: SPyNet flow estimation on a synthetic 7-frame 256x256 clip (6 pairs x 2 directions), through the drop-in
`Spynet` module (vsrlab.vsr.models.RealBasicVSR.modules.spynet.Spynet).  Prints pairs/s and the conv TFLOP/s
(479 808 flop per level-pixel, SURVEY 8d: 502.99 GFLOP per call of 12 pair-directions).

    python tools/spynet_bench.py [--clips 16]     # clips batched per call (12 pair-directions each)"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=16)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    from vsrlab.vsr.models.RealBasicVSR.modules.spynet import Spynet
    from vsrlab_b200 import functional as VF
    dev = torch.device("cuda:0")
    VF.set_precision(a.precision)
    torch.manual_seed(0)
    sp = Spynet(pretrained=False).to(dev).eval() if "pretrained" in Spynet.__init__.__code__.co_varnames else Spynet().to(dev).eval()
    clip = torch.rand(a.clips, 7, 3, 256, 256, device=dev)
    ref = torch.cat([clip[:, :-1].flatten(0, 1), clip[:, 1:].flatten(0, 1)], 0)       # backward pairs then forward pairs
    supp = torch.cat([clip[:, 1:].flatten(0, 1), clip[:, :-1].flatten(0, 1)], 0)
    with torch.no_grad():
        for _ in range(3):
            sp(ref, supp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fl = sp(ref, supp)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    pairs = ref.shape[0]
    flops = 479808.0 * (256 * 256 * 4.0 / 3.0 * (1 - 0.25 ** 6)) * pairs
    print(f"cfg2 SPyNet {a.precision}: {pairs} pair-directions of 256x256 in {ms:.2f} ms = {pairs / ms * 1e3:.0f} pairs/s, "
          f"{flops / ms / 1e9:.0f} TFLOP/s (conv FLOPs), flow mean |f| = {fl.abs().mean().item():.3f} px")


if __name__ == "__main__":
    main()
