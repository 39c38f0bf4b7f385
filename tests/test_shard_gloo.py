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


class _FakeEngine:
    """table_raw / table_adopt of gap2seq_b200.engine.Engine over host memory (32-byte slots like the device table)."""
    SLOT = 32

    def __init__(self, cap=0, n=0, fill=None):
        import numpy as np
        self.cap, self.n = cap, n
        self.buf = np.full(cap * self.SLOT, fill, dtype=np.uint8) if cap else None

    def table_raw(self):
        return self.buf.ctypes.data, self.cap, self.n, self.buf.size

    def table_adopt(self, cap, n):
        import numpy as np
        self.cap, self.n = cap, n
        self.buf = np.zeros(cap * self.SLOT, dtype=np.uint8)
        return self.buf.ctypes.data


def _bcast_worker(rank, world, port, ret):
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeEngine(1024, 400, fill=7) if rank == 0 else _FakeEngine()
    nbytes = shard.broadcast_table(eng, 0, device="cpu", as_tensor=lambda ptr, n: torch.from_numpy(eng.buf))
    ret[rank] = (nbytes, eng.cap, eng.n, int(eng.buf.sum()))
    dist.destroy_process_group()


def test_world2_table_replication_protocol():
    """broadcast_table: shape and byte count travel first, the receiver allocates and checks its own view of the buffer,
    then one broadcast of the raw bytes (the slot size is the engine's business, not the helper's)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + os.getpid() % 2000
    mp.spawn(_bcast_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] == ret[1] == (1024 * 32, 1024, 400, 7 * 1024 * 32)
