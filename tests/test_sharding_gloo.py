"""CPU tests of the multi-GPU host logic: world_size-2 gloo processes (no GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpcgpu_b200 import sharding


def test_shard_ranges_partition_the_batch():
    for batch in (0, 1, 7, 8, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.shard_range(batch, world, r)
                assert 0 <= lo <= hi <= batch
                seen += list(range(lo, hi))
                for i in range(lo, hi):
                    assert sharding.owner_of(i, batch, world) == r
            assert seen == list(range(batch))
    assert sharding.shard_range(1024, 8, 3) == (384, 512)          # BASELINE config 4: 128 per GPU
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(batch, world, rank)
        # flag of system i is (i % 3 == 0): every rank must end up with the same global vector
        local = torch.tensor([1 if i % 3 == 0 else 0 for i in range(lo, hi)], dtype=torch.uint8)
        got = sharding.gather_converged(local, batch, world, rank)
        # the pipelined form: two steps in flight, waited for in order, same answer
        p1 = sharding.gather_converged_async(local, batch, world, rank)
        p2 = sharding.gather_converged_async(1 - local, batch, world, rank)
        assert p1.wait().tolist() == got.tolist() and p2.wait().tolist() == [1 - v for v in got.tolist()]
        q.put((rank, got.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7, 1])
def test_gather_converged_world2_gloo(batch):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [1 if i % 3 == 0 else 0 for i in range(batch)]
    assert res[0] == want and res[1] == want


def test_gather_single_rank_needs_no_process_group():
    flags = torch.tensor([0, 1, 0], dtype=torch.uint8)
    assert sharding.gather_converged(flags, 3, 1, 0).tolist() == [0, 1, 0]


class _FakePlan:
    """Stand-in for solver.StepPlan on CPU (host-logic test only: there is no CPU solver in the product): marks trajectory i
    of the shard as 'hit the cap' when its first lambda entry is negative, and records what it was asked to run."""

    def __init__(self, n, m, N, batch):
        self.n, self.N, self.batch = n, N, batch
        self.flags = torch.zeros(batch, dtype=torch.uint8)
        self.calls = []

    def run(self, d_G, d_C, d_g, d_c, rho, d_lambda, d_dz, max_iter, exit_tol, direct_fallback=False):
        lam = d_lambda.view(self.batch, self.n * self.N)
        self.flags = (lam[:, 0] < 0).to(torch.uint8)
        d_dz.fill_(float(rho))
        self.calls.append((max_iter, exit_tol, direct_fallback))

    def device_flags(self):
        return self.flags


def _step_worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, m, N = 2, 1, 4
        st = sharding.ShardedStep(n, m, N, batch, world, rank, plan_factory=_FakePlan, direct_fallback=(rank == 0))
        lo, hi = sharding.shard_range(batch, world, rank)
        lam = torch.ones(max(hi - lo, 0) * n * N)
        for j, i in enumerate(range(lo, hi)):
            if i % 2 == 1:
                lam[j * n * N] = -1.0                     # odd global trajectories "do not converge"
        dz = torch.zeros(max(hi - lo, 0) * ((n + m) * (N - 1) + n))
        flags = st.step(None, None, None, None, 0.5, lam, dz, 7, 1e-3)
        q.put((rank, flags.tolist(), st.local, st.plan.calls if st.plan else None, float(dz[0]) if dz.numel() else None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [6, 5, 1])
def test_sharded_step_world2_gloo(batch):
    """ShardedStep host logic: every rank runs its shard through its plan and all ranks end with the same global flag vector
    (one all-gather), including a ragged tail and a rank that owns nothing."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_step_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r[0]: r[1:] for r in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [i % 2 for i in range(batch)]
    assert res[0][0] == want and res[1][0] == want
    assert res[0][1] + res[1][1] == batch
    assert res[0][2] == [(7, 1e-3, True)]                 # rank 0 asked for the direct fallback
    if res[1][1]:
        assert res[1][2] == [(7, 1e-3, False)] and res[1][3] == 0.5
    else:
        assert res[1][2] is None
