"""World-size-2 ``gloo`` tests (CPU) of the data-parallel host logic: group sharding, the flat SUM all-reduce and the
global-normaliser rule.  The kernels themselves need a B200; what is covered here is everything around them."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reactranker_b200.parallel import GradSync, broadcast_parameters, shard_groups, shard_rows


def test_shard_groups_cover_and_balance():
    rng = np.random.default_rng(0)
    for G, world in ((82, 8), (8, 8), (128, 4), (5, 2), (3, 8), (1, 2)):
        atoms = rng.integers(12, 29, size=G) * rng.integers(20, 60, size=G)
        runs = shard_groups(atoms, world)
        assert len(runs) == world and runs[0][0] == 0 and runs[-1][1] == G
        assert all(a[1] == b[0] for a, b in zip(runs, runs[1:])) and all(lo <= hi for lo, hi in runs)
        if G >= world:
            assert all(hi > lo for lo, hi in runs)
            loads = [atoms[lo:hi].sum() for lo, hi in runs]
            assert max(loads) <= atoms.sum() / world + atoms.max()          # within one group of the ideal share
    assert shard_rows([3, 4, 5, 6], 1, 3) == (3, 12)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)                               # different init per rank ...
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 1))
        broadcast_parameters(model)                           # ... made equal
        w0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        # a ListNet-style loss: mean over ALL items of the global batch = sum over ranks of (local sum / global N)
        g = torch.Generator().manual_seed(7)
        X, t = torch.randn(12, 6, generator=g), torch.randn(12, generator=g)
        lo, hi = [(0, 5), (5, 12)][rank]
        local = ((model(X[lo:hi]).squeeze(-1) - t[lo:hi]) ** 2).sum() / 12
        local.backward()
        GradSync(model.parameters())()
        flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        out[rank] = (w0.numpy(), flat.numpy())
    finally:
        dist.destroy_process_group()


def test_gloo_world2_flat_allreduce_equals_single_process():
    world = 2
    port = 29500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        (w_a, g_a), (w_b, g_b) = out[0], out[1]
    assert np.array_equal(w_a, w_b) and np.allclose(g_a, g_b)
    # single-process reference on the whole batch with rank 0's weights
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 1))
    g = torch.Generator().manual_seed(7)
    X, t = torch.randn(12, 6, generator=g), torch.randn(12, generator=g)
    ((model(X).squeeze(-1) - t) ** 2).mean().backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).numpy()
    assert np.allclose(g_a, want, rtol=1e-5, atol=1e-7)
