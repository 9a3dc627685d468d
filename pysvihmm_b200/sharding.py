"""Data parallelism over meta-observations: the windows of one minibatch are independent given
the global parameters (the reference only *adds* their results, hmmsgd_metaobs.py:430-433), so
rank r takes windows r, r+G, r+2G, ... and ONE sum all-reduce of the packed statistics per global
step replaces the accumulation.  Every rank then applies the natural-gradient step redundantly on
bitwise-identical inputs, so replicas stay in lock-step without a broadcast."""
import numpy as np


def dist_or_none():
    """torch.distributed when initialised with world_size > 1, else None."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
    except Exception:
        pass
    return None


def shard_starts(starts, rank, world):
    """This rank's share of the window start indices (strided, sizes differ by at most 1)."""
    return np.ascontiguousarray(np.asarray(starts, dtype=np.int64)[rank::world])


def allreduce_stats(stats, dist=None):
    """Sum the packed statistics tensor (include/svihmm.h layout) over ranks, in place."""
    dist = dist_or_none() if dist is None else dist
    if dist is not None:
        dist.all_reduce(stats)
    return stats
