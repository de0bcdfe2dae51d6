"""Chunk -> GPU sharding (no collective on the data path: chunks are independent).

``shard_indices`` is the round-robin rule (arrival index modulo n_gpus); ``shard_of_chunk_id`` is the
equivalent rule on the chunk's uuid.  ``max_over_ranks`` is the only cross-rank exchange the benchmark
needs: the per-rank elapsed time reduced with MAX (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import List


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_items, world))


def shard_of_chunk_id(chunk_id: str, world: int) -> int:
    return int(chunk_id, 16) % world


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar over the default process group (returns value itself when not initialised)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
