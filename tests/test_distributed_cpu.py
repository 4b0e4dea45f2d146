"""Host-side multi-rank logic on CPU: world_size-2 gloo process group, contiguous scenario shards
and the final gather/merge of per-rank summaries (the only collective of the path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from model_predictive_control_b200 import distributed as D


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 8, 1 << 20, 8388608 + 3):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)


def test_merge_summaries():
    rows = torch.tensor([[4, 10.0, 0.5, 3, 1, 0, 40, 3], [6, 5.0, 0.7, 1, 0, 2, 70, 4]], dtype=torch.float64)
    m = D.merge_summaries(rows)
    assert m == {"scenarios": 10.0, "sum_cost": 15.0, "max_violation": 0.7, "sum_saturated": 4.0, "n_infeasible": 1.0,
                 "n_max_iter": 2.0, "sum_iters": 110.0, "n_solved": 7.0}


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = D.shard_range(total, rank, world)
        cost = torch.arange(lo, hi, dtype=torch.float64)          # each rank's own shard of per-scenario costs
        summ = torch.tensor([hi - lo, cost.sum(), float(rank + 1) * 0.25, hi - lo, rank, 0, 11 * (hi - lo), hi - lo - rank],
                            dtype=torch.float64)
        merged = D.gather_summaries(summ)
        if rank == 0:
            q.put(merged)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    total = 1001
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    merged = q.get()
    assert merged["scenarios"] == total
    assert merged["sum_cost"] == total * (total - 1) / 2
    assert merged["max_violation"] == 0.5 and merged["n_infeasible"] == 1 and merged["sum_iters"] == 11 * total
