"""Data-parallel plumbing for N replicas (one process per GPU): request sharding and the max-over-ranks timing rule.

The hot path shards across independent sequences only: rank r owns requests {i : i mod N == r},
runs them on its own full-weight replica with no data-path collective, and the job throughput is
(sum of units over ranks) / (max time over ranks).  `torch.distributed` (nccl on GPUs, gloo in the CPU tests) carries only
these scalars.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n_items: int, world: int, rank: int) -> list[int]:
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_items, world))


def job_throughput(units_local: float, seconds_local: float, device=None) -> tuple[float, float, float]:
    """-> (whole-job units/s, total units, max seconds).  Works without an initialised process group (N=1)."""
    t = torch.tensor([units_local, seconds_local], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        u = t[:1].clone()
        s = t[1:].clone()
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        dist.all_reduce(s, op=dist.ReduceOp.MAX)
        units, secs = float(u.item()), float(s.item())
    else:
        units, secs = float(t[0]), float(t[1])
    return units / secs, units, secs


def gather_tokens(tokens_local: torch.Tensor) -> torch.Tensor:
    """All-gather of the per-rank greedy token ids of one decode step ([b_local] -> [world * b_local], rank-major)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tokens_local
    out = [torch.empty_like(tokens_local) for _ in range(dist.get_world_size())]
    dist.all_gather(out, tokens_local)
    return torch.cat(out)
