"""Sharding of independent gaps over the GPUs of one box (the reference parallelises per record with
a thread pool, src/Gap2Seq.cpp:296-323).  One process per GPU; no data-path collective: the k-mer table is
replicated by a single broadcast, gaps are assigned longest-processing-time-first by estimated cost, and the
per-gap results are gathered in input order on every rank (the traceback must then run in input order on ONE rank,
because it consumes libc rand() sequentially)."""
from __future__ import annotations

import heapq


def estimate_cost(gap_len: int, gap_err: int, fuz: int, width: float = 1.0) -> float:
    """Levels of the DP (g + e + 2 fuz) times the expected border width."""
    return (gap_len + gap_err + 2 * fuz) * width


def lpt_shards(costs: list[float], n_ranks: int) -> list[list[int]]:
    """Longest-processing-time-first: returns, per rank, the indices of its gaps (each list in input order)."""
    heap = [(0.0, r) for r in range(n_ranks)]
    heapq.heapify(heap)
    shards: list[list[int]] = [[] for _ in range(n_ranks)]
    for i in sorted(range(len(costs)), key=lambda j: (-costs[j], j)):
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + costs[i], r))
    return [sorted(s) for s in shards]


def merge_in_input_order(per_rank: list[list], shards: list[list[int]]) -> list:
    """Inverse of the sharding: per_rank[r][j] is the result of gap shards[r][j]."""
    n = sum(len(s) for s in shards)
    out = [None] * n
    for res, idx in zip(per_rank, shards):
        assert len(res) == len(idx)
        for v, i in zip(res, idx):
            out[i] = v
    return out


def gather_results(local: list, shards: list[list[int]]):
    """all_gather_object of the per-rank result lists (gloo or nccl process group), merged in input order."""
    import torch.distributed as dist
    world = dist.get_world_size()
    bucket = [None] * world
    dist.all_gather_object(bucket, local)
    return merge_in_input_order(bucket, shards)


class DeviceBytes:
    """Zero-copy view of a raw device buffer for torch (``torch.as_tensor(DeviceBytes(ptr, n), device='cuda')``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def broadcast_table(engine, src: int = 0, device: str = "cuda", as_tensor=None):
    """Replicate rank `src`'s k-mer table to every rank's engine with ONE NCCL broadcast (NVLink / NVSwitch).
    `device` / `as_tensor(ptr, nbytes)` exist for the CPU test of the protocol (gloo, host buffers)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    if as_tensor is None:
        as_tensor = lambda ptr, nbytes: torch.as_tensor(DeviceBytes(ptr, nbytes), device="cuda")   # noqa: E731
    meta = torch.zeros(3, dtype=torch.int64, device=device)
    if rank == src:
        ptr, cap, n, nbytes = engine.table_raw()
        meta[0], meta[1], meta[2] = cap, n, nbytes
    dist.broadcast(meta, src=src)
    cap, n, nbytes = int(meta[0].item()), int(meta[1].item()), int(meta[2].item())
    if rank != src:
        ptr = engine.table_adopt(cap, n)
        _, cap2, _, nbytes2 = engine.table_raw()   # the receiver's own view of the buffer it just allocated
        if (cap2, nbytes2) != (cap, nbytes):
            raise RuntimeError(f"table replication: receiver expects {nbytes2} bytes for {cap2} slots, sender has {nbytes} for {cap}")
    t = as_tensor(ptr, nbytes)
    dist.broadcast(t, src=src)
    if device == "cuda":
        torch.cuda.synchronize()
    return nbytes
