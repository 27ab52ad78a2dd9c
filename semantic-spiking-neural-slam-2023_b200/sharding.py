"""Trial sharding across GPUs.

Trials are independent, so the data path has no collective: rank ``r`` of ``W`` owns a contiguous
block of the global trial ids, builds the same network (one seed -> identical static weights on
every GPU) and steps only its block.  The only exchange is one ``all_gather`` of a few per-trial
error statistics after the run (NCCL on GPUs; the same code runs over gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(n_trials, rank, world):
    """Contiguous block [lo, hi) of global trial ids owned by ``rank`` (sizes differ by at most one)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_trials), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def trial_seeds(n_trials, rank, world, base_seed=0):
    """Per-trial seeds of this rank's block: a trial's start state depends on its GLOBAL id only."""
    lo, hi = shard_range(n_trials, rank, world)
    return [base_seed + i for i in range(lo, hi)]


def gather_trial_stats(local, n_trials, device=None):
    """all_gather per-trial statistics ``[n_local, k]`` -> ``[n_trials, k]`` in global trial order.

    Blocks may be ragged (``n_trials`` not divisible by the world size): every rank pads to the
    largest block, and the padding is dropped after the collective."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(np.atleast_2d(np.asarray(local, dtype=np.float32)))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        if local.shape[0] != n_trials:
            raise ValueError("single-process gather: local block must hold every trial")
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_range(n_trials, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} owns {hi - lo} trials but passed {local.shape[0]} rows")
    width = -(-n_trials // world)
    buf = torch.zeros((width, local.shape[1]), dtype=torch.float32)
    buf[:local.shape[0]] = torch.from_numpy(local)
    if device is not None:
        buf = buf.to(device)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    parts = []
    for r, t in enumerate(outs):
        a, b = shard_range(n_trials, r, world)
        parts.append(t[:b - a].cpu().numpy())
    return np.concatenate(parts, axis=0)


def max_over_ranks(value, device=None):
    """Max of a scalar over ranks (timing is reported as the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
