"""Partition of independent links over ranks (one process per GPU) and the bench's reductions.

The reference scatters contiguous row chunks from MPI rank 0 to ranks 1..size-1
(main.cpp:275-307: baseChunk = n / workers, the first `remainder` workers get one more row) and never
gathers: each rank writes its own files (main.cpp:796-797).  Links are independent (no equation reads
next_stream), so the partition needs no data-path collective; the only collectives are the
bench's MAX over ranks of the device time and SUM of the step counts.
"""
from __future__ import annotations


def shard_range(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Rows [lo, hi) of rank `rank` under the reference's chunk rule (main.cpp:275-307)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_timing(ms_local: float, sums_local, dist=None, device=None):
    """(max over ranks of ms, element-wise sum over ranks of sums).  dist=None: single process."""
    import torch
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    s = torch.tensor([float(x) for x in sums_local], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t.item()), [float(x) for x in s.tolist()]
