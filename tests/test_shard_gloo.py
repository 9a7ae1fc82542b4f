"""Sequence sharding: pure partitioning properties, and a world_size-2 gloo run on CPU
that checks union-of-shards == single-process result (no data-path collective; only the
final gather communicates)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tcsfm_b200 import shard


def test_shard_ranges_tile_exactly():
    for n in (0, 1, 5, 133, 795, 1000):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_kitti_seq09_schedule():
    mbs = shard.window_minibatches(1591, stride=2, minibatch=6)
    assert sum(len(m) for m in mbs) == 795 and len(mbs) == 133 and len(mbs[-1]) == 3
    assert mbs[0] == [1, 3, 5, 7, 9, 11]


def _window_result(centre):
    # stand-in for one optimised window: a deterministic [6] "pose" that depends only on the window
    g = torch.Generator().manual_seed(centre)
    return torch.randn(6, generator=g)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mbs = shard.window_minibatches(61, stride=2, minibatch=6)
    lo, hi = shard.shard_range(len(mbs), rank, world)
    local = torch.stack([_window_result(c) for mb in mbs[lo:hi] for c in mb]) if hi > lo else torch.zeros(0, 6)
    full = shard.gather_results(local, world)
    if rank == 0:
        torch.save(full, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, port, out_path), nprocs=2, join=True)
    gathered = torch.load(out_path)
    mbs = shard.window_minibatches(61, stride=2, minibatch=6)
    single = torch.stack([_window_result(c) for mb in mbs for c in mb])
    assert torch.equal(gathered, single)
