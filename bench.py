"""Benchmark of the Real-BasicVSR x4 hot path (BASELINE.json: output frames/s, 720p out).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one `model(lr)` pass of the drop-in `vsrlab...RealBasicVSR` over `--clips`
synthetic 30-frame 180x320 clips per GPU (cfg3 of BASELINE.json, experiment=basic 5/5 blocks
unless --blocks is given), bf16 mode.  Prints ONE JSON line (rank 0).

* value        frames/s, all ranks, inputs already resident in HBM (CUDA events, max over ranks)
* e2e          same metric through the public nn.Module call with pinned HOST buffers: H2D of the
               clip and D2H of sr/lq inside the timed region
* roofline     the dominant kernel (tcgen05 implicit-GEMM conv): algorithmic conv FLOPs / CUDA-event
               time of those launches, against MEASURED_PEAKS.json
* cpu_baseline the CPU oracle (restatement of the reference's PyTorch path, oracle/) timed on this
               box's host cores on a bounded sample
* --impl reference: the reference's CPU path (oracle port; /root/reference cannot travel to the
               GPU box) on all host cores, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "realbasicvsr_x4_output_frames_per_sec_720p"
UNIT = "frames/s"
T_FRAMES, LR_H, LR_W = 30, 180, 320


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sust": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "src": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sust": 1400.0, "src": "fallback"}


def build_model(blocks: int, device):
    from vsrlab.vsr.models.RealBasicVSR.realbasicvsr import RealBasicVSR
    torch.manual_seed(0)
    m = RealBasicVSR(cleaning_blocks=blocks, mid_channels=64, upscale=4, res_blocks=blocks, pretrained_flow=False,
                     train_flow=False)
    return m.to(device).eval()


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent (NVML) clocks and throttle reasons during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        self.power_w, self.power_limit_w = [], None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown"}
            try:
                self.power_limit_w = nv.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0
            except Exception:  # noqa: BLE001
                pass
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.power_w.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:  # noqa: BLE001
                    pass
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.05)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        sm = sorted(self.sm)
        pw = sorted(self.power_w)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "power_w": pw[len(pw) // 2] if pw else None, "power_limit_w": self.power_limit_w}


def bind_to_gpu_numa_node(index: int):
    """Run this process on the CPUs NVML reports as local to GPU `index`.  The end-to-end number moves ~0.7 GB per step
    between pinned host memory and the GPU; with the host buffers on the far NUMA node the same run measured 40 % lower."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to gpu {index}"
    except Exception as e:  # noqa: BLE001
        return f"unchanged ({type(e).__name__})"
    return "unchanged"


def oracle_sample(blocks: int, frames: int, threads: int):
    """CPU restatement of the reference path on `frames` frames of one 180x320 clip."""
    from oracle import vsr_oracle as O
    torch.set_num_threads(threads)
    sd = {k: v.detach().cpu() for k, v in build_model(blocks, "cpu").state_dict().items()}
    x = torch.rand(1, frames, 3, LR_H, LR_W, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        t0 = time.perf_counter()
        O.realbasicvsr(x, sd)
        dt = time.perf_counter() - t0
    return frames / dt, dt


def run_reference(a):
    """--impl reference: the reference's CPU implementation (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = 3
    for _ in range(a.warmup):
        oracle_sample(a.blocks, 2, cores)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        oracle_sample(a.blocks, frames, cores)
    dt = time.perf_counter() - t0
    v = a.steps * frames / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(a, 1),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{frames} frames of one 180x320 clip per step, full Real-BasicVSR forward (oracle/vsr_oracle.py)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(a, clips):
    return {"workload": f"cfg3: Real-BasicVSR x4 inference, {T_FRAMES}-frame {LR_H}x{LR_W}->{4*LR_H}x{4*LR_W} clips, "
                        f"{a.blocks}/{a.blocks} blocks", "clips_per_gpu_per_step": clips, "frames_per_clip": T_FRAMES,
            "precision": "bf16 activations, fp32 accumulate", "cuda_graph": "whole forward captured once, replayed per step", "parallelism": f"clips sharded over {a.gpus} GPU(s), no collective",
            "l2": "per-step working set (GBs of activations) exceeds the 126 MB L2; no explicit flush"}


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else a library prints to stdout while the
    benchmark runs (e.g. NCCL's version banner) has been redirected to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--blocks", type=int, default=5)
    ap.add_argument("--clips", type=int, default=2, help="clips per GPU per step")
    ap.add_argument("--streams", type=int, default=1,
                    help="process the step's clips on this many CUDA streams (fills the fill/drain bubbles of the "
                         "sequential propagation kernels with another clip's work)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp32_ffma"],
                    help="bf16 (headline) | fp32 = fp32-accurate split-bf16 on the tensor cores | fp32_ffma = FFMA kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl != "reference" else a.warmup
    if a.impl == "reference":
        return run_reference(a)

    import torch.distributed as dist
    from vsrlab_b200 import functional as VF
    from vsrlab_b200 import ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)       # pinned host buffers are then first-touched next to the GPU's PCIe root
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    VF.set_precision(a.precision)
    model = build_model(a.blocks, dev)
    clips = a.clips
    gen = torch.Generator().manual_seed(1000 + rank)
    host_lr = torch.rand(clips, T_FRAMES, 3, LR_H, LR_W, generator=gen).pin_memory()
    lr_dev = host_lr.to(dev)
    work = torch.empty_like(lr_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    side = [torch.cuda.Stream(device=dev) for _ in range(max(0, a.streams - 1))]

    def step():
        work.copy_(lr_dev)                       # the model refines its input in place (reference contract)
        if a.streams <= 1:
            with torch.no_grad():
                return model(work)
        # clips are independent: split them over the streams (same public call, one per stream)
        cur = torch.cuda.current_stream(dev)
        parts = work.chunk(a.streams, dim=0)
        outs = []
        for i, part in enumerate(parts):
            st = cur if i == 0 else side[i - 1]
            st.wait_stream(cur) if i else None
            with torch.cuda.stream(st), torch.no_grad():
                outs.append(model(part))
        for st in side:
            cur.wait_stream(st)
        return outs

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0, r0 = ops.launch_count(), VF.replayed_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count() - l0 + VF.replayed_launches() - r0      # direct launches + kernel nodes of graph replays
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- end to end: pinned host -> device -> model -> pinned host --------------------------
    # What a user of the drop-in does for a stream of clips: H2D of clip k+1 and D2H of clip k-1 run on a copy
    # stream while `model(lr)` of clip k runs on the compute stream.  Every step's H2D and D2H are inside the
    # timed region; the region ends when the last result has landed in pinned host memory.
    host_sr = [torch.empty(clips, T_FRAMES, 3, 4 * LR_H, 4 * LR_W).pin_memory() for _ in range(2)]
    host_lq = [torch.empty_like(host_lr).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)

    def e2e_run(n_steps):
        pending = None                                    # (sr, lq, done_event) of the previous step
        nxt = None
        with torch.cuda.stream(copy_stream):
            nxt = host_lr.to(dev, non_blocking=True)
            up = torch.cuda.Event()
            up.record(copy_stream)
        for k in range(n_steps):
            x, ready = nxt, up
            main_stream.wait_event(ready)
            x.record_stream(main_stream)
            with torch.no_grad():
                sr, lq = model(x)
            done = torch.cuda.Event()
            done.record(main_stream)
            with torch.cuda.stream(copy_stream):
                if k + 1 < n_steps:                       # prefetch the next clip batch
                    nxt = host_lr.to(dev, non_blocking=True)
                    up = torch.cuda.Event()
                    up.record(copy_stream)
                copy_stream.wait_event(done)
                sr.record_stream(copy_stream)
                lq.record_stream(copy_stream)
                host_sr[k % 2].copy_(sr, non_blocking=True)
                host_lq[k % 2].copy_(lq, non_blocking=True)
        main_stream.wait_stream(copy_stream)

    e2e_run(4)        # warm-up long enough for the caching allocator to own every output block the steady state needs
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_run(a.steps)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- per-kernel roofline: CUDA events around every launch of one more step --------------
    # The timed region replays a CUDA graph, so the per-launch pass must not be paced by the host either: a spin kernel
    # holds the stream while the step's launches and event records are queued, then the GPU runs them back to back and the
    # events see device-side durations (kernel + the GPU's own launch latency), not waits for Python.
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.15 * 1.9e9))
    ops.PROFILE = []
    step()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    fam, parts = {}, {}
    shapes = {}
    for kind, ev0, ev1, work_units, tag in prof:
        sec = ev0.elapsed_time(ev1) * 1e-3
        if "[" in kind:                                   # VSRB_PROFILE_SHAPES=1: per-shape table, family key without the shape
            q = shapes.setdefault(tag + " " + kind, [0.0, 0.0, 0])
            q[0] += sec
            q[1] += work_units
            q[2] += 1
            kind = kind.split("[")[0]
        d = fam.setdefault(kind, [0.0, 0.0, 0])
        d[0] += sec
        d[1] += work_units
        d[2] += 1
        if kind.startswith("conv_"):
            q = parts.setdefault(tag, [0.0, 0.0, 0])
            q[0] += sec
            q[1] += work_units
            q[2] += 1
    step_s_prof = sum(d[0] for d in fam.values())

    # ---- flow_warp on a working set far larger than L2 (256 feature maps, 1.9 GB in + 1.9 GB out), L2 flushed ----
    def warp_standalone():
        from vsrlab_b200._lib import BF16, PAD_ZEROS
        n = 256
        xw = torch.randn(n, LR_H, LR_W, 64, device=dev).to(torch.bfloat16)
        fw = (torch.rand(n, LR_H, LR_W, 2, device=dev) - 0.5) * 4.0
        ow = torch.empty_like(xw)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for i in range(8):
            flush.zero_()
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            ops.flow_warp(xw, fw, ow, n, LR_H, LR_W, 64, BF16, PAD_ZEROS)
            w1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(w0.elapsed_time(w1) * 1e-3)
        return n * LR_H * LR_W * 264 / (sorted(ts)[len(ts) // 2]) / 1e9

    warp_big = warp_standalone() if rank == 0 else 0.0

    if ops.debug_status() != 0:                     # a kernel gave up on a pipeline barrier: the numbers would be meaningless
        raise RuntimeError("libvsrb200 reported a pipeline time-out during the benchmark")
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    frames_step = world * clips * T_FRAMES
    value = frames_step * a.steps / (ms * 1e-3)
    e2e = frames_step * a.steps / (ms_e2e * 1e-3)
    if rank == 0:
        pk = peaks()
        traffic = {}
        tp = ROOT / "profiles" / "traffic.json"            # dram bytes per launch from `ncu --set full` captures
        if tp.exists():
            traffic = json.loads(tp.read_text())
        conv = fam.get("conv_tc") or fam.get("conv_tc_x3") or fam.get("conv_f32") or [1e-9, 0.0, 0]
        warp = fam.get("flow_warp", [1e-9, 0.0, 0])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp32": "bf16x3 (split-bf16, fp32-accurate)", "fp32_ffma": "f32"}[a.precision],
            "data": "synthetic", "config": workload_config(a, clips),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": host_lr.numel() * 4,
                    "d2h_bytes_per_step": (host_sr[0].numel() + host_lq[0].numel()) * 4,
                    "pipeline": "copies on a second stream overlap the next step's compute; all copies are inside the timed region",
                    "host_affinity": numa},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": {"kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, all launches of a step)", "bound": "tensor",
                         "achieved": conv[1] / conv[0] / 1e12, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                         "frac": conv[1] / conv[0] / 1e12 / pk["tf_sust"],
                         "traffic": (traffic.get("conv_tc") or {}).get("dram_bytes_per_launch"),
                         "traffic_source": traffic.get("conv_tc"),
                         "peak_source": pk["src"] + " (sustained)",
                         "launches_per_step": conv[2], "share_of_step": conv[0] / max(step_s_prof, 1e-9)},
            "roofline_warp": {"kernel": "flow_warp_kernel", "bound": "hbm", "achieved": warp[1] / warp[0] / 1e9, "peak": pk["hbm"],
                              "unit": "GB/s", "frac": warp[1] / warp[0] / 1e9 / pk["hbm"],
                              "traffic": (traffic.get("flow_warp") or {}).get("dram_bytes_per_launch"), "traffic_source": traffic.get("flow_warp"),
                              "launches_per_step": warp[2], "share_of_step": warp[0] / max(step_s_prof, 1e-9),
                              "note": "in-step calls move 30 MB each (L2 resident, launch-latency bound); `standalone` is the "
                                      "same kernel on 256 feature maps (3.9 GB moved, L2 flushed)",
                              "standalone": {"achieved": warp_big, "frac": warp_big / pk["hbm"], "unit": "GB/s"}},
            "kernel_time_share": {k: round(v[0] / max(step_s_prof, 1e-9), 4) for k, v in fam.items()},
            "conv_by_part": {k: {"ms": round(v[0] * 1e3, 3), "tflops": round(v[1] / v[0] / 1e12, 1), "launches": v[2]}
                             for k, v in parts.items()},
        }
        if shapes:
            line["conv_by_shape"] = {k: {"ms": round(v[0] * 1e3, 3), "tflops": round(v[1] / v[0] / 1e12, 1), "launches": v[2]}
                                     for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][0])}
        if world == 1 and not a.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, dtc = oracle_sample(a.blocks, 3, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"3 frames of one 180x320 clip ({dtc:.1f} s), full forward, oracle/vsr_oracle.py"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
