"""Latency of one small clip (BASELINE cfg1: 1x5x3x64x64) with and without CUDA-graph replay."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vsrlab.vsr.models.RealBasicVSR.realbasicvsr import RealBasicVSR
from vsrlab_b200 import functional as VF
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = RealBasicVSR(cleaning_blocks=5, mid_channels=64, upscale=4, res_blocks=5, pretrained_flow=False, train_flow=False).to(dev).eval()
x = torch.rand(1, 5, 3, 64, 64, device=dev)
for graphs in (False, True):
    VF.GRAPHS = graphs
    for mode in ("bf16", "fp32"):
        with torch.no_grad(), VF.precision(mode):
            for _ in range(3):
                net(x.clone())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                net(x.clone())
            torch.cuda.synchronize()
            print(f"graphs={graphs} {mode}: {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms per 5-frame 64x64 clip", flush=True)
