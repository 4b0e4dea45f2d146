"""Scenario sharding across the GPUs of one box and the final gather of summaries.

Scenarios are independent: rank r of W owns the contiguous shard ``shard_range(total, r, W)`` and
nothing is exchanged inside a solve.  The only collective is the final gather of one 8-number
summary per rank (``torch.distributed`` all_gather: NCCL over NVLink on the GPUs, gloo in the CPU
tests), merged by :func:`merge_summaries`.
"""
from __future__ import annotations

from ctypes import POINTER, c_double, c_int, c_int32, c_int64, c_void_p

import torch

from . import _lib

_lib.register("mpc_summary", c_int, [c_void_p] * 5 + [c_int64, c_void_p, c_int, c_void_p])

FIELDS = ("scenarios", "sum_cost", "max_violation", "sum_saturated", "n_infeasible", "n_max_iter", "sum_iters",
          "n_solved")


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of ``total`` scenarios for ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_summary(cost=None, violation=None, n_saturated=None, status=None, iters=None, batch=None):
    """Device reduction (K6) of per-scenario results to the 8-number summary (float64 tensor)."""
    ref = next(t for t in (cost, violation, n_saturated, status, iters) if t is not None)
    _lib.require_cuda(ref)
    n = batch if batch is not None else ref.numel()
    out = torch.empty(8, dtype=torch.float64, device=ref.device)
    fl = cost if cost is not None else violation
    dt = _lib.dtype_enum(fl) if fl is not None else _lib.MPC_F64
    with torch.cuda.device(ref.device):
        _lib.check(_lib.lib().mpc_summary(_lib.ptr(cost), _lib.ptr(violation), _lib.ptr(n_saturated), _lib.ptr(status),
                                          _lib.ptr(iters), n, _lib.ptr(out), dt, _lib.stream(ref.device)))
    return out


def merge_summaries(rows: torch.Tensor) -> dict:
    """rows [world, 8] -> totals (sums, except the maximum violation)."""
    rows = rows.double()
    tot = rows.sum(dim=0)
    tot[2] = rows[:, 2].max()
    return {k: float(v) for k, v in zip(FIELDS, tot.tolist())}


def gather_summaries(summary: torch.Tensor) -> dict:
    """All-gather the per-rank summary and merge; without an initialised process group: one rank."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts = [torch.empty_like(summary) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, summary)
        return merge_summaries(torch.stack(parts).cpu())
    return merge_summaries(summary[None].cpu())
