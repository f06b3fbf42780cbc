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
