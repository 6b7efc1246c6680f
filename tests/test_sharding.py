"""CPU tests of the N>1 plumbing with the gloo backend, world_size 2: batch
partitioning, max-over-ranks timing, whole-job throughput, result gather -- and a two-rank run of
the host layer (native calls emulated) whose gathered outputs equal the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tepose_b200 import shard


def test_partition_covers_range_exactly():
    for n in (0, 1, 7, 32, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard.partition(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.partition(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import synth
        from tests import fake_native
        from tests.helpers import build_product_model
        torch.set_num_threads(1)
        B, T = 5, 3
        lo, hi = shard.partition(B, world, rank)
        x = torch.from_numpy(synth.make_input(9, B, T))
        model, _ = build_product_model(9, T, 1, 32)
        with fake_native.install():
            out = model(x[lo:hi])[-1]
        verts = shard.gather_rows(out["verts"], B)
        theta = shard.gather_rows(out["theta"], B)
        fps, ms = shard.aggregate_throughput(hi - lo, 10.0 * (rank + 1))
        mx = shard.max_over_ranks([1.0 + rank, 5.0 - rank])
        if rank == 0:
            q.put(dict(verts=verts.numpy(), theta=theta.numpy(), fps=fps, ms=ms, mx=mx))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_forward_matches_single_process():
    from oracle import synth
    from tests import fake_native
    from tests.helpers import build_product_model
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    B, T = 5, 3
    x = torch.from_numpy(synth.make_input(9, B, T))
    model, _ = build_product_model(9, T, 1, 32)
    with fake_native.install():
        full = model(x)[-1]
    np.testing.assert_allclose(res["verts"], full["verts"].numpy(), atol=1e-6)
    np.testing.assert_allclose(res["theta"], full["theta"].numpy(), atol=1e-6)
    assert res["ms"] == 20.0 and abs(res["fps"] - B / 20e-3) < 1e-6      # all units / slowest rank
    assert res["mx"] == [2.0, 5.0]
