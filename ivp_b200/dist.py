"""Host-side sharding helpers for the one-process-per-GPU launch (torchrun) and the multi-device context.

Trajectories are independent units: the ensemble is split into contiguous shards, no data-path collective
exists.  torch.distributed is only used for the barrier and for reducing the measured time (MAX)
and the step counts (SUM) over ranks.
"""
from __future__ import annotations

import os
from typing import Tuple


def dist_env() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process => (0, 0, 1))."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_range(N: int, g: int, G: int) -> Tuple[int, int]:
    """Static contiguous split [g*N/G, (g+1)*N/G) -- the same arithmetic libivpb uses across its devices."""
    return N * g // G, N * (g + 1) // G


def weak_offset(n_per_rank: int, rank: int) -> int:
    """Global index of a rank's first trajectory under weak scaling (fixed work per GPU)."""
    return n_per_rank * rank


def reduce_time_and_count(ms: float, count: float, device=None, use_dist: bool = False) -> Tuple[float, float]:
    """MAX over ranks of the device time, SUM over ranks of the processed units."""
    if not use_dist:
        return float(ms), float(count)
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    c = torch.tensor([count], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), float(c.item())
