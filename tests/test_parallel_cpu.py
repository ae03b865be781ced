"""World-size-2 ``gloo`` tests (CPU) of the data-parallel host logic: group sharding, the flat SUM all-reduce and the
global-normaliser rule.  The kernels themselves need a B200; what is covered here is everything around them."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reactranker_b200.parallel import GradSync, broadcast_parameters, init_from_env, plan_shard, shard_groups, shard_rows


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


def test_plan_shard_cuts_whole_groups_and_covers_every_row():
    rng = np.random.default_rng(1)
    for G, world in ((82, 8), (8, 8), (5, 2), (3, 8)):
        scope = rng.integers(2, 60, size=G)
        atoms = rng.integers(12, 29, size=int(scope.sum()))
        off = np.concatenate(([0], np.cumsum(scope)))
        parts = [plan_shard(scope, atoms, r, world) for r in range(world)]
        assert parts[0][2] == 0 and parts[-1][3] == scope.sum()
        for (gl, gh, rl, rh), nxt in zip(parts, parts[1:] + [None]):
            assert (rl, rh) == (off[gl], off[gh])                               # row range == the groups' rows: no group is split
            if nxt is not None:
                assert (gh, rh) == (nxt[0], nxt[2])


def test_dp_normalisers_select_global_groups_or_items():
    from reactranker_b200.train import loss as RL
    assert RL.MLEloss()._norm(5) == 5 and RL.ListnetLoss()._norm(50, items=True) == 50 and RL._items_norm(50) == 50.0
    with RL.dp_normalisers(groups=16, items=400):
        assert RL.MLEloss()._norm(5) == 16 and RL.evidential_ranking()._norm(5) == 16 and RL.MLEDisLoss()._norm(5) == 16
        assert RL.ListnetLoss()._norm(50, items=True) == 400 and RL._items_norm(50) == 400.0
        assert RL.MLEloss(global_norm=3)._norm(5) == 3                          # an explicit constructor argument wins
        with RL.dp_normalisers(groups=2, items=9):
            assert RL.MLEloss()._norm(5) == 2
        assert RL.MLEloss()._norm(5) == 16
    assert RL.MLEloss()._norm(5) == 5


class _FlatModel(torch.nn.Module):
    """Stands in for ReactionModel's gradient layout on the CPU: backward() leaves p.grad as views of one flat buffer."""

    def __init__(self):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 1))
        self._grad_flat, self._grad_offsets = None, None

    def hot_parameters(self):
        return list(self.net.parameters())

    def alias_grads(self):
        ps = self.hot_parameters()
        offs, total = [], 0
        for p in ps:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        flat = torch.zeros(total)
        for p, o in zip(ps, offs):
            flat[o:o + p.numel()] = p.grad.reshape(-1)
            p.grad = flat[o:o + p.numel()].view(p.shape)
        self._grad_flat, self._grad_offsets = flat, offs


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    assert init_from_env(device=None) == (rank, world)                         # the entry points' way into the group (gloo without a GPU)
    try:
        # fast path: gradients aliased into one flat buffer are all-reduced in place
        torch.manual_seed(rank)
        fm = _FlatModel()
        broadcast_parameters(fm)
        g = torch.Generator().manual_seed(7)
        X, t = torch.randn(12, 6, generator=g), torch.randn(12, generator=g)
        lo, hi = [(0, 5), (5, 12)][rank]
        (((fm.net(X[lo:hi]).squeeze(-1) - t[lo:hi]) ** 2).sum() / 12).backward()
        fm.alias_grads()
        sync = GradSync(fm.hot_parameters(), None, fm)
        sync()
        assert (sync.fast_path_steps, sync.copy_path_steps) == (1, 0)
        out[("fast", rank)] = torch.cat([p.grad.reshape(-1) for p in fm.hot_parameters()]).numpy()
        # mixed paths in ONE collective: rank 0 all-reduces its aliased buffer in place, rank 1 had an empty shard (no backward pass,
        # p.grad is None) and goes through the staging bucket -- same layout, same element count, or NCCL / gloo would mismatch
        torch.manual_seed(0)
        mm = _FlatModel()
        sm = GradSync(mm.hot_parameters(), None, mm)
        if rank == 0:
            ((mm.net(X).squeeze(-1) - t) ** 2).mean().backward()
            mm.alias_grads()
        sm()
        assert (sm.fast_path_steps, sm.copy_path_steps) == ((1, 0) if rank == 0 else (0, 1))
        out[("mixed", rank)] = torch.cat([p.grad.reshape(-1) for p in mm.hot_parameters()]).numpy()
        # copy path with a NON-CONTIGUOUS gradient and a missing one (empty shard): the reduced values must land in p.grad itself
        w = torch.nn.Parameter(torch.zeros(3, 4))
        w.grad = torch.full((4, 3), float(rank + 1)).t()                        # strides (1, 3)
        b = torch.nn.Parameter(torch.zeros(2))
        if rank == 0:
            b.grad = torch.ones(2)
        s2 = GradSync([w, b])
        s2()
        assert s2.copy_path_steps == 1 and not w.grad.is_contiguous()
        out[("copy", rank)] = (w.grad.clone().numpy(), b.grad.numpy())
    finally:
        dist.destroy_process_group()


def _worker_old(rank, world, port, out):
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


def test_gloo_world2_inplace_flat_allreduce_and_copy_path():
    world = 2
    port = 31500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert np.array_equal(res[("fast", 0)], res[("fast", 1)])
    torch.manual_seed(0)
    fm = _FlatModel()
    g = torch.Generator().manual_seed(7)
    X, t = torch.randn(12, 6, generator=g), torch.randn(12, generator=g)
    ((fm.net(X).squeeze(-1) - t) ** 2).mean().backward()
    want = torch.cat([p.grad.reshape(-1) for p in fm.hot_parameters()]).numpy()
    assert np.allclose(res[("fast", 0)], want, rtol=1e-5, atol=1e-7)
    for r in (0, 1):
        assert np.array_equal(res[("copy", r)][0], np.full((3, 4), 3.0)) and np.array_equal(res[("copy", r)][1], np.ones(2))
    # rank 1 contributed zeros: both ranks hold rank 0's whole-batch gradient
    assert np.allclose(res[("mixed", 0)], want, rtol=1e-5, atol=1e-7) and np.array_equal(res[("mixed", 0)], res[("mixed", 1)])


def test_gloo_world2_flat_allreduce_equals_single_process():
    world = 2
    port = 29500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_old, args=(world, port, out), nprocs=world, join=True)
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
