"""N>1 host logic on CPU: world_size-2 gloo process group, clip sharding and the throughput
reduction bench.py uses (sum of frames / max of seconds).  No data-path collective exists."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vsrlab_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_clips(7, rank, world)
    done = shard.run_sharded(lambda cid: cid * 10, list(range(100, 107)), rank, world)
    frames, secs = shard.reduce_throughput(len(mine) * 30, 1.0 + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, mine, done, frames, secs, gathered))
    dist.destroy_process_group()


def test_clip_sharding_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, d0, f0, s0, g0), (r1, m1, d1, f1, s1, g1) = res
    assert m0 == [0, 2, 4, 6] and m1 == [1, 3, 5]
    assert sorted(m0 + m1) == list(range(7))                     # every clip exactly once
    assert d0 == [(100, 1000), (102, 1020), (104, 1040), (106, 1060)]
    assert f0 == f1 == 7 * 30 and s0 == s1 == 2.0                # sum of frames, max of seconds
    assert g0 == g1 == [m0, m1]


def test_single_process_identity_and_windows():
    assert shard.reduce_throughput(60, 1.5) == (60, 1.5)
    assert shard.shard_clips(5, 0, 1) == [0, 1, 2, 3, 4]
    assert shard.shard_clips(2, 3, 4) == []                      # ragged: more ranks than clips
    assert shard.windows(70, 32) == [(0, 32), (32, 64), (64, 70)]
