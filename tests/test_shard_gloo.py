"""N>1 host logic on CPU: two processes (gloo) shard one clip, the union is exact and disjoint, and the
bench's max-over-ranks reduction behaves."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, sizes, q):
    sys.path.insert(0, ROOT)
    from motioncam_decoder_b200 import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = {
        "contiguous": list(shard.shard_contiguous(n_frames, world, rank)),
        "round_robin": list(shard.shard_round_robin(n_frames, world, rank)),
        "bytes": list(shard.shard_by_bytes(sizes, world)[rank]),
        "weak": list(shard.weak_scaling_clip(7, world, rank)),
    }
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # the bench's timing reduction: every rank reports its own elapsed ms, the job's time is the max
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((gathered, float(t.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [0, 1, 5, 240])
def test_two_ranks_partition_the_clip(n_frames):
    world = 2
    sizes = [1000 + 37 * (i % 11) + (5000 if i > n_frames // 2 else 0) for i in range(n_frames)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, sizes, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 11.0
    for mode in ("contiguous", "round_robin", "bytes"):
        parts = [g[mode] for g in gathered]
        flat = sorted(i for part in parts for i in part)
        assert flat == list(range(n_frames)), mode
    assert abs(len(gathered[0]["contiguous"]) - len(gathered[1]["contiguous"])) <= 1
    assert sorted(gathered[0]["weak"] + gathered[1]["weak"]) == list(range(14))
    if n_frames >= 5:
        b0 = sum(sizes[i] for i in gathered[0]["bytes"])
        b1 = sum(sizes[i] for i in gathered[1]["bytes"])
        assert abs(b0 - b1) <= max(sizes), (b0, b1)


def test_shard_edge_cases():
    sys.path.insert(0, ROOT)
    from motioncam_decoder_b200 import shard
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 8, 9, 1000):
            parts = [list(shard.shard_contiguous(n, world, r)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            parts = shard.shard_by_bytes([3] * n, world)
            assert [i for p in parts for i in p] == list(range(n))
    with pytest.raises(ValueError):
        shard.shard_contiguous(4, 2, 2)
