"""N>1 host logic on CPU: world_size-2 `gloo` process group (127.0.0.1 rendezvous)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "chainer-faster-rcnn_b200"))
    from frcnn_b200 import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert shard.env_rank() == (rank, rank, world)
        mine = shard.shard_indices(7, rank, world)
        t = shard.max_over_ranks(1.0 + rank)                       # slowest rank decides
        counts = shard.gather_ints(100 + rank)
        thr = shard.aggregate_throughput(10, 1.0 + rank)
        # the training path's one collective: SUM of the flat gradient bucket, identical on every rank afterwards
        g = torch.arange(6, dtype=torch.float32) * (rank + 1)
        shard.allreduce_sum_(g)
        assert torch.equal(g, torch.arange(6, dtype=torch.float32) * 3)
        dist.barrier()
        q.put((rank, mine, t, counts, thr, shard.image_seed(rank, 3)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world, port = 2, 29533
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, m0, t0, c0, thr0, s0), (r1, m1, t1, c1, thr1, s1) = res
    assert m0 == [0, 2, 4, 6] and m1 == [1, 3, 5]                 # disjoint cover, image i -> rank i mod N
    assert t0 == t1 == 2.0 and c0 == c1 == [100, 101]
    assert thr0 == thr1 == pytest.approx(2 * 10 / 2.0)            # all items / slowest rank
    assert s0 != s1


def test_single_process_fallbacks():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "chainer-faster-rcnn_b200"))
    from frcnn_b200 import shard
    assert shard.max_over_ranks(3.5) == 3.5 and shard.gather_ints(7) == [7]
    g = torch.ones(3)
    assert shard.allreduce_sum_(g) is g and torch.equal(g, torch.ones(3))
    assert shard.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        shard.shard_indices(5, 2, 2)
