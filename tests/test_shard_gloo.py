"""CPU tests (-m "not gpu") of the N>1 path: shard plan and the final counter reduction, world_size 2, gloo."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from pomcpp_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = shard.shard_plan(rank, world, 1000)
    counters = np.arange(10, dtype=np.int64) * (rank + 1)          # rank 0: k, rank 1: 2k
    total = shard.reduce_counters(counters, dist)
    t = shard.max_over_ranks([1.0 + rank, 5.0 - rank], dist)
    q.put((rank, plan, total.tolist(), t.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_and_counter_reduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    (r0, plan0, tot0, t0), (r1, plan1, tot1, t1) = out
    assert plan0["first"] == 0 and plan1["first"] == 1000 and plan0["count"] == plan1["count"] == 1000
    assert plan0["total"] == 2000
    assert tot0 == tot1 == [3 * k for k in range(10)]
    assert t0 == t1 == [2.0, 5.0]


def test_single_rank_passthrough():
    from pomcpp_b200 import shard
    c = np.arange(10, dtype=np.int64)
    assert (shard.reduce_counters(c) == c).all()
    assert shard.shard_plan(0, 1, 7) == {"first": 0, "count": 7, "total": 7}
    with pytest.raises(ValueError):
        shard.shard_plan(2, 2, 7)


def test_strong_plan_partitions_a_fixed_total():
    from pomcpp_b200 import shard
    for total, world in ((1 << 20, 8), (1000, 3), (5, 8)):
        plans = [shard.strong_plan(r, world, total) for r in range(world)]
        assert plans[0]["first"] == 0 and sum(p["count"] for p in plans) == total
        for a, b in zip(plans, plans[1:]):
            assert a["first"] + a["count"] == b["first"]
        assert max(p["count"] for p in plans) - min(p["count"] for p in plans) <= 1
