"""Sharding a parameter sweep over GPUs: one process per GPU, block-cyclic by parameter index.

Sweep points are independent ODEs, so there is no collective on the data path:
rank r of `world` integrates rows r, r + world, r + 2*world, ... of the parameter table and the results
are scattered back to their global positions (by the caller, or with `gather_results` through
torch.distributed when every rank needs them).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_indices(n_points: int, world: int, rank: int) -> np.ndarray:
    """Global indices of the sweep points owned by `rank` (block-cyclic, block = 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, n_points, world)


def shard_params(params: np.ndarray, world: int, rank: int) -> Tuple[np.ndarray, np.ndarray]:
    idx = shard_indices(len(params), world, rank)
    return np.ascontiguousarray(np.asarray(params)[idx]), idx


def scatter_back(global_out: np.ndarray, local_out: np.ndarray, idx: np.ndarray) -> None:
    """Place a rank's results at their global sweep positions (disjoint slices: no reduction)."""
    global_out[idx] = local_out


def gather_results(local_out: np.ndarray, n_points: int, group=None) -> Optional[np.ndarray]:
    """All ranks contribute their shard; every rank returns the full [n_points, ...] array.
    Uses torch.distributed (gloo on CPU tensors, nccl on CUDA tensors) only for this final gather."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (n_points + world - 1) // world
    t = torch.from_numpy(np.ascontiguousarray(local_out))
    is_complex = t.is_complex()
    if is_complex:
        t = torch.view_as_real(t)
    pad_shape = (per,) + tuple(t.shape[1:])
    buf = torch.zeros(pad_shape, dtype=t.dtype)
    buf[: t.shape[0]] = t
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    full = torch.zeros((n_points,) + tuple(t.shape[1:]), dtype=t.dtype)
    for r in range(world):
        idx = shard_indices(n_points, world, r)
        full[torch.from_numpy(idx)] = bufs[r][: len(idx)]
    if is_complex:
        full = torch.view_as_complex(full)
    return full.numpy()
