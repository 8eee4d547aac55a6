"""Test helper for the N > 1 (one process per GPU) deployment: every rank verifies its contiguous slice and the ranks
exchange the status bytes with one all_gather (gloo on CPU).  Test infrastructure only: the product has no collective."""
from __future__ import annotations

from typing import Callable

import numpy as np

from jubjub_schnorr_b200.sharding import shard_range


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
