"""Multi-GPU plumbing of the sampler: one process per GPU (torchrun), disjoint Philox
index ranges per rank, and ONE exchange -- the all-reduce (sum, int64) of the
bmc_tallies buffer.  Records stay on the GPU that produced them.

Because every raw deal is a pure function of (seed, index), the union of the ranks'
ranges -- hence every tally -- is identical for any number of ranks.
"""
from __future__ import annotations

import os

STEP_STRIDE_LOG2 = 40          # index space reserved per (step, rank) in open-ended (accept-target) jobs


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def rank_range(first_index: int, n_raw: int, rank: int, world: int):
    """Contiguous share of [first_index, first_index + n_raw) for `rank`: (first, count).
    The first `n_raw % world` ranks take one extra index, so the shares tile the range exactly."""
    base, extra = divmod(int(n_raw), world)
    count = base + (1 if rank < extra else 0)
    start = first_index + rank * base + min(rank, extra)
    return start, count


def step_first_index(step: int, rank: int, world: int) -> int:
    """First index of an accept-target job of `step` on `rank`: ranges of 2^40 indices that never overlap."""
    return (step * world + rank) << STEP_STRIDE_LOG2


def init_process_group(backend: str | None = None, device=None):
    """torch.distributed init from the torchrun environment (nccl on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    rank, world, _ = env_rank_world()
    if world == 1 or dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl" and device is not None:
        dist.init_process_group(backend, device_id=device)
    else:
        dist.init_process_group(backend)


def allreduce_tallies(tallies):
    """Sum the int64 tally tensor (bmc_tallies words) over all ranks, in place; no-op for one rank."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tallies, op=dist.ReduceOp.SUM)
    return tallies


def max_over_ranks(values):
    """Element-wise max of a float64 tensor over ranks (device timings are reported as the slowest rank's)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(values, op=dist.ReduceOp.MAX)
    return values
