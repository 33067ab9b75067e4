"""Chunk sharding + host merge (the N>1 path), on CPU: pure logic, and a world_size-2 gloo run in which each
rank scores its own shards (with the oracle standing in for the GPU -- tests may use it) and rank 0 merges."""
import os
import socket
import sys

import numpy as np
import pytest

from blamm_b200 import shard
from blamm_b200.capi import HIT_DTYPE
from oracle import oracle as O
from tests import util


def test_plan_covers_stream_exactly_once():
    for n, world, halo, chunk in [(0, 2, 5, 10), (1, 2, 5, 10), (100, 3, 7, 10), (101, 8, 34, 33), (10, 4, 20, 3)]:
        shards = shard.plan_shards(n, world, halo, chunk)
        assert sum(s.n_payload for s in shards) == n
        pos = 0
        for k, s in enumerate(shards):
            assert s.start == pos and s.rank == k % world and s.index == k
            assert s.n_payload <= chunk and s.n_total == min(s.n_payload + halo, n - s.start)
            pos += s.n_payload
    with pytest.raises(ValueError):
        shard.plan_shards(10, 0, 1, 1)


def _scan_shards(case, shards, rank):
    out = []
    for s in shards:
        if s.rank != rank:
            continue
        block = bytes(case["chars"][s.start:s.start + s.n_total])
        fs = np.concatenate([np.zeros(1, np.uint64), shard.local_frag_starts(case["frag_start"], s)])
        pos, col, sc = O.scan_stream(block, fs, case["P"], case["col_len"], case["thr"], n_payload=s.n_payload)
        h = np.zeros(len(pos), dtype=HIT_DTYPE)
        h["pos"], h["col"], h["score"] = pos, col, sc
        out.append((s.index, h))
    return out


def test_sharded_equals_single_pass():
    case = util.random_case(5, n_motifs=10, n_nt=60_000)
    halo = int(case["col_len"].max()) - 1
    for world, chunk in [(1, 60_000), (2, 7001), (8, 1000), (3, 59_999)]:
        shards = shard.plan_shards(len(case["chars"]), world, halo, chunk)
        parts = {}
        for r in range(world):
            parts.update(dict(_scan_shards(case, shards, r)))
        merged = shard.merge_hits([parts[s.index] for s in shards], shards)
        pos, col, sc = O.scan_stream(bytes(case["chars"]), case["frag_start"], case["P"], case["col_len"], case["thr"])
        assert np.array_equal(merged["pos"], pos) and np.array_equal(merged["col"], col)
        assert np.array_equal(merged["score"].view(np.uint32), sc.view(np.uint32))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = util.random_case(6, n_motifs=8, n_nt=40_000)
    halo = int(case["col_len"].max()) - 1
    shards = shard.plan_shards(len(case["chars"]), world, halo, 5003)
    mine = _scan_shards(case, shards, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)       # host-side merge only: no collective on the data path itself
    if rank == 0:
        parts = dict(kv for g in gathered for kv in g)
        merged = shard.merge_hits([parts[s.index] for s in shards], shards)
        pos, col, sc = O.scan_stream(bytes(case["chars"]), case["frag_start"], case["P"], case["col_len"], case["thr"])
        q.put(bool(np.array_equal(merged["pos"], pos) and np.array_equal(merged["col"], col)
                   and np.array_equal(merged["score"].view(np.uint32), sc.view(np.uint32)) and len(pos) > 0))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world_size_2():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_chunk_ordered_concatenation_is_the_merge():
    """What bench.py and the CLI rely on since the device returns every chunk's hits in (position, column) order: the merge in
    reference order is the plain concatenation of the per-chunk lists in chunk order (positions + the chunk's start) -- no sort,
    whatever rank scored which chunk."""
    case = util.random_case(9, n_motifs=12, n_nt=90_000)
    halo = int(case["col_len"].max()) - 1
    pos, col, sc = O.scan_stream(bytes(case["chars"]), case["frag_start"], case["P"], case["col_len"], case["thr"])
    for world, chunk in [(2, 9001), (8, 4096), (3, 90_000)]:
        shards = shard.plan_shards(len(case["chars"]), world, halo, chunk)
        per_rank = {r: dict(_scan_shards(case, shards, r)) for r in range(world)}
        cat_pos, cat_col, cat_sc = [], [], []
        for s in shards:                                              # chunk order, not rank order
            h = per_rank[s.rank][s.index]
            assert np.all(np.diff(h["pos"].astype(np.int64)) >= 0)    # (the oracle, like the device, lists a chunk in position order)
            cat_pos.append(h["pos"] + np.uint64(s.start)); cat_col.append(h["col"]); cat_sc.append(h["score"])
        assert np.array_equal(np.concatenate(cat_pos), pos) and np.array_equal(np.concatenate(cat_col), col)
        assert np.array_equal(np.concatenate(cat_sc).view(np.uint32), sc.view(np.uint32))
