"""Multi-GPU plumbing: reads shard across ranks (one process per GPU, full index on each); the only
collective of the path is the sum of the five hit counters of HitSink (hit.h:169-175)."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of read ids owned by `rank`: rank r gets [r*n/W, (r+1)*n/W)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def counters_from_found(found: np.ndarray, khits: int, mhits: int, all_hits: bool) -> np.ndarray:
    """{aligned, unaligned, maxed, reported, reportedPaired} for one shard, from hitsForThisRead_ per read
    (HitSinkPerThread::finishRead + HitSink::tallyAlignments, hit.h:741-786, 214-223)."""
    lim = 0xFFFFFFFF if all_hits else khits
    maxed = found > mhits
    nrep = np.where(maxed, 0, np.minimum(found, lim)).astype(np.int64)
    aligned = int(np.count_nonzero(nrep))
    nmax = int(np.count_nonzero(maxed))
    return np.array([aligned, len(found) - aligned - nmax, nmax, int(nrep.sum()), 0], dtype=np.int64)


def allreduce_counters(counters: np.ndarray, device=None) -> np.ndarray:
    """Sum the five counters over all ranks (NCCL on GPUs, gloo in CPU tests); identity when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counters
    t = torch.tensor(np.asarray(counters), dtype=torch.int64, device=device if device is not None else "cpu")   # a copy: the caller's array is left alone
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
