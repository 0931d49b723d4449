"""Multi-GPU plumbing of the hot path: envs are independent, so they shard contiguously over ranks with no
data-path collective; the only exchange is the final reduction of the episode counters (and, for timing,
a MAX over ranks).  torch.distributed is used as plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_plan(rank, world, envs_per_gpu):
    """Global env range [first, first + count) owned by `rank` (weak scaling: fixed work per GPU)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return {"first": rank * envs_per_gpu, "count": envs_per_gpu, "total": world * envs_per_gpu}


def strong_plan(rank, world, total):
    """Global env range owned by `rank` when a FIXED total is split over the ranks (strong scaling): contiguous ranges,
    sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return {"first": first, "count": base + (1 if rank < extra else 0), "total": total}


def reduce_counters(counters, dist=None, device="cpu"):
    """Sum of the POM_STATS_WORDS episode counters over all ranks (the one collective of a run)."""
    c = np.ascontiguousarray(counters, dtype=np.int64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return c.copy()
    import torch
    t = torch.from_numpy(c.copy()).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def max_over_ranks(values, dist=None, device="cpu"):
    """Element-wise MAX over ranks of a few float64 timings (a multi-GPU number is the slowest rank's)."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return v.copy()
    import torch
    t = torch.from_numpy(v.copy()).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()
