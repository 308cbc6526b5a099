"""Sharding of Monte-Carlo samples / sweep points over ranks (one process per GPU).

The path has no data-path collective: sample i belongs to rank i // per_rank (contiguous
blocks), every rank advances its own batch, and the only communication is a gather of the result
waveforms and per-sample statistics to rank 0 (NCCL over NVLink on the GPUs, gloo in the CPU
tests)."""
import numpy as np


def shard(n_total, rank, world):
    """contiguous [lo, hi) block of the samples owned by `rank`"""
    per = (n_total + world - 1) // world
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


def gather_results(dist, tensor, dst=0):
    """gather equally-shaped per-rank result tensors to `dst`; returns the concatenation on dst"""
    import torch
    world = dist.get_world_size()
    rank = dist.get_rank()
    lst = [torch.empty_like(tensor) for _ in range(world)] if rank == dst else None
    dist.gather(tensor, lst, dst=dst)
    return torch.cat(lst, dim=0) if rank == dst else None
