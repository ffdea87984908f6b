"""Multi-GPU plumbing of the hot path: batch-sharded replicas, no data-path collective.

Every (batch, head) pair is an independent recurrence, so N GPUs run N
independent shards of the batch axis; the only communication is the timing reduction of the
benchmark (max over ranks) and, in training, DDP's gradient all-reduce which belongs to the
caller (ultralytics/engine/trainer.py:277), not to this op.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def batch_shard(global_batch: int, rank: int, world: int):
    """Contiguous shard [start, start+size) of the batch axis for `rank`; sizes differ by at most one."""
    base, rem = divmod(global_batch, world)
    size = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, size


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (e.g. milliseconds measured with CUDA events on each rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def job_throughput(units_this_rank: float, ms_this_rank: float, device=None) -> float:
    """Whole-job units per second: all ranks' units over the slowest rank's time."""
    total = sum_over_ranks(units_this_rank, device)
    ms = max_over_ranks(ms_this_rank, device)
    return total / (ms * 1e-3)
