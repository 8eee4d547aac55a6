"""Contiguous sharding of one batch over ranks (one process per GPU under torchrun), the partition of SURVEY.md section 8(e).

Items are independent: there is no exchange step and no collective on the data path.  bench.py uses shard_range to cut the
2^24-item strong-scaling batch (BASELINE.json configs[4]) into per-rank slices; within one process the same contiguous
partition over the devices of a context is done by the library itself (csrc/kernels.cu, plan_shards)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's contiguous slice; slices differ by at most one item and cover [0, n) exactly."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
