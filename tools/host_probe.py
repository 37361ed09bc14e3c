"""What limits device->host copies when all ranks of one box copy at once?  Run under torchrun; every rank reports its CPU
affinity, the memory nodes it may use, its GPU's NUMA node, and its D2H / H2D GB/s alone and with all ranks copying
concurrently, for pinned buffers allocated (a) as is, (b) after binding to the GPU-local CPUs.
    python -m torch.distributed.run --nproc-per-node 8 tools/host_probe.py
"""
import os
import subprocess
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def bw(dst, src, stream, reps=4):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
    stream.synchronize()
    dt = time.perf_counter() - t0
    return src.numel() * src.element_size() * reps / dt / 1e9


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    info = {"rank": rank, "cpus_allowed": len(os.sched_getaffinity(0)), "cpu_count": os.cpu_count()}
    try:
        st = Path("/proc/self/status").read_text()
        info["mems_allowed"] = [ln.split(":")[1].strip() for ln in st.splitlines() if ln.startswith("Mems_allowed_list")][0]
        info["cpus_allowed_list"] = [ln.split(":")[1].strip() for ln in st.splitlines() if ln.startswith("Cpus_allowed_list")][0]
    except Exception as e:  # noqa: BLE001
        info["status_err"] = str(e)
    nodes = sorted(p.name for p in Path("/sys/devices/system/node").glob("node[0-9]*")) if Path("/sys/devices/system/node").exists() else []
    info["numa_nodes"] = nodes
    n = 176 * 1024 * 1024                       # 704 MB of fp32, one step's sr + lq
    gpu = torch.empty(n, device=dev)
    s = torch.cuda.Stream()
    host_a = torch.empty(n).pin_memory()
    res = {"default": {}}
    res["default"]["d2h_all"] = bw(host_a, gpu, s)
    res["default"]["h2d_all"] = bw(gpu, host_a, s)
    for r in range(world):                      # one rank at a time
        if r == rank:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.cuda.stream(s):
                for _ in range(4):
                    host_a.copy_(gpu, non_blocking=True)
            s.synchronize()
            res["default"]["d2h_alone"] = n * 4 * 4 / (time.perf_counter() - t0) / 1e9
        dist.barrier()
    del host_a
    import bench
    info["bind"] = bench.bind_to_gpu_numa_node(local)
    host_b = torch.empty(n).pin_memory()
    res["bound"] = {"d2h_all": bw(host_b, gpu, s), "h2d_all": bw(gpu, host_b, s)}
    # half of the ranks at a time (0-3, then 4-7): is it a per-socket limit?
    for half in range(2):
        if (rank * 2) // world == half:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.cuda.stream(s):
                for _ in range(4):
                    host_b.copy_(gpu, non_blocking=True)
            s.synchronize()
            res["bound"]["d2h_half"] = n * 4 * 4 / (time.perf_counter() - t0) / 1e9
        dist.barrier()
    # uint8 payload (a quarter of the bytes)
    gpu8 = torch.empty(n, dtype=torch.uint8, device=dev)
    host8 = torch.empty(n, dtype=torch.uint8).pin_memory()
    res["bound"]["d2h_all_u8_GBps"] = bw(host8, gpu8, s, reps=16)
    out = [None] * world
    dist.all_gather_object(out, (info, res))
    if rank == 0:
        for i, r_ in out:
            print(i)
            print("   ", r_)
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
        except Exception as e:  # noqa: BLE001
            print("topo:", e)
        tot = lambda k1, k2: sum(r_[k1][k2] for _, r_ in out)  # noqa: E731
        print("aggregate GB/s: default d2h_all", tot("default", "d2h_all"), "bound d2h_all", tot("bound", "d2h_all"), "bound h2d_all",
              tot("bound", "h2d_all"), "u8", tot("bound", "d2h_all_u8_GBps"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
