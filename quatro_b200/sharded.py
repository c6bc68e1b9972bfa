"""Multi-GPU sharding of a batch of independent registration pairs.

One process per GPU (torch.distributed).  Pairs are independent end to end, so there is NO data-path
collective: rank r registers the pairs {p : p mod world == r} on its own handle, and the only
communication is one all_gather of the fixed-size per-pair result records (NCCL over NVLink on GPUs;
gloo in the CPU tests of the host logic).  A pair's result is bit-identical whichever rank runs it.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

from .capi import RESULT_DTYPE


def shard_indices(n_pairs: int, rank: int, world: int) -> np.ndarray:
    """Global pair ids owned by `rank` (static round-robin: equal counts up to +-1)."""
    return np.arange(rank, n_pairs, world, dtype=np.int64)


def gather_results(local: np.ndarray, n_pairs: int, rank: int, world: int, device=None) -> np.ndarray:
    """all_gather the per-rank RESULT_DTYPE records and return all n_pairs records in global order."""
    assert local.dtype == RESULT_DTYPE and len(local) == len(shard_indices(n_pairs, rank, world))
    if world == 1:
        return local.copy()
    import torch
    import torch.distributed as dist
    cap = (n_pairs + world - 1) // world
    buf = np.zeros(cap, RESULT_DTYPE)
    buf[: len(local)] = local
    t = torch.from_numpy(buf.view(np.uint8).reshape(cap, -1).copy())
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = np.zeros(n_pairs, RESULT_DTYPE)
    for r, o in enumerate(outs):
        idx = shard_indices(n_pairs, r, world)
        rec = np.ascontiguousarray(o.cpu().numpy()).view(RESULT_DTYPE).reshape(-1)
        full[idx] = rec[: len(idx)]
    return full


def register_sharded(register_local: Callable[[Sequence[int]], np.ndarray], n_pairs: int, rank: int, world: int, device=None) -> np.ndarray:
    """register_local(ids) -> RESULT_DTYPE array for the given global pair ids (this rank's GPU)."""
    ids = shard_indices(n_pairs, rank, world)
    local = register_local(ids) if len(ids) else np.zeros(0, RESULT_DTYPE)
    return gather_results(local, n_pairs, rank, world, device)
