"""Frame-parallel sharding of a clip across the GPUs of one box.

Frames are independent (every frame buffer is self-contained and loadFrame is random access by timestamp,
/root/reference/lib/Decoder.cpp:184-201), so multi-GPU decode is a partition of the timestamp-sorted frame list
with no exchange step: each rank owns a context on its own GPU, decodes its frames, keeps its outputs.  Nothing
here talks to a GPU; bench.py and the feeders call these helpers, tests/test_shard_gloo.py checks them across
two processes.
"""
from typing import List, Sequence


def shard_contiguous(n_frames: int, world: int, rank: int) -> range:
    """Contiguous chunk of the frame list for `rank` (file-read locality); sizes differ by at most one."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def shard_round_robin(n_frames: int, world: int, rank: int) -> range:
    """Frames rank, rank + world, ...: balances clips whose block statistics drift over time."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    return range(rank, n_frames, world)


def shard_by_bytes(sizes: Sequence[int], world: int) -> List[range]:
    """Contiguous partition of frames into `world` chunks with near-equal compressed bytes (greedy on the prefix
    sums): the decode time of a frame follows its algorithmic bytes, not its index."""
    if world <= 0:
        raise ValueError("bad world")
    n = len(sizes)
    total = float(sum(sizes))
    out, start, acc = [], 0, 0.0
    for r in range(world):
        if r == world - 1:
            end = n
        else:
            target = total * (r + 1) / world
            end = start
            while end < n - (world - 1 - r) and acc + sizes[end] / 2.0 <= target:
                acc += sizes[end]
                end += 1
        out.append(range(start, end))
        start = end
    return out


def weak_scaling_clip(frames_per_gpu: int, world: int, rank: int) -> range:
    """Global frame indices of `rank` when every GPU decodes `frames_per_gpu` frames of one long clip (bench.py)."""
    return shard_contiguous(frames_per_gpu * world, world, rank)
