"""Multi-GPU plumbing for the hot path: sequences (or bodies) are independent, so the batch axis is
cut contiguously across ranks and NO data-path collective is needed.  The only
communication is the timing reduction of the benchmark (max over ranks) and an optional gather of
results on the caller's side."""
from __future__ import annotations

import torch
import torch.distributed as dist


def partition(n: int, world: int, rank: int):
    """Contiguous, balanced shard [lo, hi) of n independent units for `rank` of `world`
    (the first n % world ranks take one extra unit)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world / rank")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (device-timed numbers are reported as
    the max over ranks).  Works with NCCL (device tensor) and gloo (CPU tensor)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def aggregate_throughput(local_units: int, local_ms: float, device=None):
    """Whole-job throughput = units processed by ALL ranks / max-over-ranks time."""
    total = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    ms = max_over_ranks([local_ms], device)[0]
    return float(total[0]) / (ms * 1e-3), ms


def gather_rows(local: torch.Tensor, n_total: int):
    """All-gather variable-length row shards back into [n_total, ...] order (caller-side
    convenience; not on the hot path)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [partition(n_total, world, r) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
