"""Partitioning of independent work units (PFT window minibatches, training minibatches)
across one-process-per-GPU ranks.  The hot path has no data-path collective (SURVEY.md
§8e): every window deep-copies the network and builds a fresh optimiser (reference
optimization_experiments/optimizer.py:177-182,211-214), so shards never talk to each
other; the only (optional) communication is the final gather of the per-window results.
"""
import os

import torch


def shard_range(n_items, rank, world_size):
    """Contiguous [start, stop) of `n_items` for `rank`; sizes differ by at most one and the
    ranges of all ranks tile [0, n_items) exactly."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def window_minibatches(n_frames, stride=2, minibatch=6):
    """The reference's PFT schedule: sliding 3-frame windows (n_frames - 2 of them,
    data/kitti_loader_stereo.py:214-223), every `stride`-th window
    (run_sequential_optimization.py:108), grouped into minibatches of `minibatch`.
    Returns a list of lists of centre-frame indices (KITTI seq 09: 1591 frames -> 795
    windows -> 133 minibatches, the last one of 3)."""
    centres = list(range(1, n_frames - 1))[::stride]
    return [centres[i:i + minibatch] for i in range(0, len(centres), minibatch)]


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def gather_results(local, world_size, group=None):
    """All-gathers per-rank result tensors of different first-dimension length (the [n,6]
    pose arrays / per-window losses) and returns them concatenated in rank order.  Works with
    any torch.distributed backend; a no-op without an initialised process group."""
    import torch.distributed as dist
    if world_size == 1 or not dist.is_available() or not dist.is_initialized():
        return local
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world_size)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s) for s in sizes]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    out = [torch.zeros_like(buf) for _ in range(world_size)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:n] for o, n in zip(out, sizes)], 0)
