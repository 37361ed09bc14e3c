"""Multi-GPU partitioning of the inference path: clips are independent units (a clip's
recurrence is sequential in time and never split), so rank r of W takes clips r, r+W, ...
and the data path needs NO collective.  The only communication is the final reduction of
(frames processed, seconds) used to report whole-job throughput.  Reference: the reference
shards data with DistributedSampler (core/utils.py:199-200) and runs inference windows
independently (test.py:125-131)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_clips(num_clips: int, rank: int, world: int) -> List[int]:
    """Indices of the clips rank `rank` of `world` processes (round-robin, like DistributedSampler
    without shuffling or padding)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, num_clips, world))


def windows(num_frames: int, window: int) -> List[Tuple[int, int]]:
    """Independent temporal windows of one video, as reference test.py:125-131 cuts them."""
    return [(i, min(i + window, num_frames)) for i in range(0, num_frames, window)]


def reduce_throughput(frames: int, seconds: float, device=None) -> Tuple[int, float]:
    """(sum of frames over ranks, max of seconds over ranks); identity without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return frames, seconds
    t = torch.tensor([float(frames)], dtype=torch.float64, device=device)
    s = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(s, op=dist.ReduceOp.MAX)
    return int(t.item()), float(s.item())


def run_sharded(process_clip, clip_ids: Sequence[int], rank: int, world: int) -> List[Tuple[int, object]]:
    """Apply `process_clip(clip_id)` to this rank's share; returns [(clip_id, result)]."""
    return [(i, process_clip(i)) for i in shard_clips(len(clip_ids), rank, world) for i in [clip_ids[i]]]
