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


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tepose_b200.train import GradSync
        torch.manual_seed(100 + rank)
        names = ["regressor.fc2.weight", "regressor.fc2.bias", "encoder.gru_fwd.weight_hh_l0", "encoder.gru_fwd.weight_ih_l0"]
        shapes = [(8, 8), (8,), (12, 4), (12, 7)]
        params = {n: torch.nn.Parameter(torch.zeros(s)) for n, s in zip(names, shapes)}
        grads = {n: torch.randn(s) for n, s in zip(names, shapes)}
        sync = GradSync()
        # buckets become ready in backward order (regressor first, recurrent weights, input weights last), as in train.py
        sync.ready([(n, grads[n]) for n in names[:2]])
        sync.ready([(names[2], grads[names[2]])])
        sync.ready([(names[3], grads[names[3]])])
        for n in names:                       # what autograd's AccumulateGrad leaves behind: the local gradient
            params[n].grad = grads[n].clone()
        sync.finish(params)
        if rank == 0:
            q.put({n: params[n].grad.numpy() for n in names} | {"bytes": sync.bytes})
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_buckets_average_across_ranks():
    """Data-parallel training plumbing (BASELINE configs[4]): three buckets all-reduced in the order the backward completes
    them; every rank ends with the mean of the per-rank gradients in param.grad."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    names = ["regressor.fc2.weight", "regressor.fc2.bias", "encoder.gru_fwd.weight_hh_l0", "encoder.gru_fwd.weight_ih_l0"]
    shapes = [(8, 8), (8,), (12, 4), (12, 7)]
    expect = {}
    for rank in range(world):
        torch.manual_seed(100 + rank)
        for n, s in zip(names, shapes):
            expect[n] = expect.get(n, 0) + torch.randn(s) / world
    for n in names:
        np.testing.assert_allclose(res[n], expect[n].numpy(), atol=1e-6)
    assert res["bytes"] == 4 * sum(int(np.prod(s)) for s in shapes)
