"""Contiguous sharding of one batch over ranks (one process per GPU) and the host-side gather of the result bytes.

SURVEY.md section 8(e): items are independent, so the only exchange step of the whole path is collecting the
per-rank status arrays; that is a torch.distributed all_gather of N bytes (gloo on CPU, nccl on GPUs)."""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's contiguous slice; slices differ by at most one item and cover [0, n) exactly."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def verify_sharded(verify_fn: Callable, pk: np.ndarray, sig: np.ndarray, msg: np.ndarray, device=None) -> np.ndarray:
    """Every rank holds the same (pk, sig, msg) arrays, verifies its own slice with verify_fn(pk, sig, msg) -> status
    bytes, and receives the full status array.  Without an initialised process group this is a plain call."""
    import torch
    import torch.distributed as dist

    n = msg.shape[0]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(verify_fn(pk, sig, msg), dtype=np.uint8)
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(n, rank, world)
    local = np.asarray(verify_fn(pk[lo:hi], sig[lo:hi], msg[lo:hi]), dtype=np.uint8) if hi > lo else np.zeros(0, dtype=np.uint8)
    width = -(-n // world)
    dev = device if device is not None else torch.device("cpu")
    padded = torch.full((width,), 0xFF, dtype=torch.uint8, device=dev)
    padded[: hi - lo] = torch.from_numpy(local).to(dev)
    parts = [torch.empty(width, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, padded)
    out = np.empty(n, dtype=np.uint8)
    for r in range(world):
        a, b = shard_range(n, r, world)
        out[a:b] = parts[r][: b - a].cpu().numpy()
    return out
