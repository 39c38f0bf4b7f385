"""Host-side logic of the N>1 path on CPU: LPT sharding and the gather of per-gap results, world_size 2 over gloo."""
import os

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from gap2seq_b200 import shard


def test_lpt_balances_and_partitions():
    costs = [5000, 100, 100, 100, 2500, 2500, 50, 50, 4000, 1000]
    sh = shard.lpt_shards(costs, 3)
    assert sorted(i for s in sh for i in s) == list(range(len(costs)))
    loads = [sum(costs[i] for i in s) for s in sh]
    assert max(loads) <= 1.2 * sum(costs) / 3 + max(costs) * 0.0 + 400
    assert shard.lpt_shards(costs, 1) == [list(range(len(costs)))]
    assert shard.lpt_shards([], 4) == [[], [], [], []]
    assert shard.merge_in_input_order([[f"r{i}" for i in s] for s in sh], sh) == [f"r{i}" for i in range(len(costs))]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gaps = [(g, 500, 10) for g in (3000, 50, 700, 1200, 90, 2500, 60, 400, 1800)]
    costs = [shard.estimate_cost(*g) for g in gaps]
    sh = shard.lpt_shards(costs, world)
    local = [("filled", i, gaps[i][0] * 2 + 1) for i in sh[rank]]      # stand-in for the device pass of this rank's gaps
    merged = shard.gather_results(local, sh)
    ret[rank] = merged
    dist.destroy_process_group()


def test_world2_gather_in_input_order():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    want = [("filled", i, g * 2 + 1) for i, g in enumerate((3000, 50, 700, 1200, 90, 2500, 60, 400, 1800))]
    assert ret[0] == want and ret[1] == want
