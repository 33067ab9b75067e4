"""Chunk sharding of a group's filtered stream over ranks (one process per GPU) and the host-side merge.

The scan path has independent units and no exchange step: the stream is cut into
chunks of `chunk` payload characters, each extended by a halo of maxLen-1 following characters (the same
overlap rule as the reference's blocks, sequence.cpp:280-290); chunk k goes to rank k % world.  Every rank
holds the full motif matrix.  No collective touches the data path; per-rank hit lists are concatenated and
sorted by (stream position, column) on the host.
"""
from __future__ import annotations

from typing import List, NamedTuple

import numpy as np


class Shard(NamedTuple):
    index: int
    rank: int
    start: int       # stream position of the first payload character
    n_payload: int
    n_total: int     # payload + halo actually available


def plan_shards(stream_len: int, world: int, halo: int, chunk: int) -> List[Shard]:
    if chunk <= 0 or world <= 0 or halo < 0:
        raise ValueError("bad sharding parameters")
    out = []
    k = 0
    for start in range(0, stream_len, chunk):
        n_payload = min(chunk, stream_len - start)
        n_total = min(n_payload + halo, stream_len - start)
        out.append(Shard(k, k % world, start, n_payload, n_total))
        k += 1
    return out


def local_frag_starts(frag_start: np.ndarray, shard: Shard) -> np.ndarray:
    """Chunk-relative starts of the fragments that begin strictly inside the shard."""
    fs = np.asarray(frag_start, dtype=np.uint64)
    lo = np.searchsorted(fs, shard.start, side="right")
    hi = np.searchsorted(fs, shard.start + shard.n_total, side="left")
    return fs[lo:hi] - np.uint64(shard.start)


def merge_hits(per_shard_hits: List[np.ndarray], shards: List[Shard]) -> np.ndarray:
    """Hits with chunk-relative positions -> one array with stream positions, in (position, column) order."""
    parts = []
    for h, s in zip(per_shard_hits, shards):
        g = h.copy()
        g["pos"] += np.uint64(s.start)
        parts.append(g)
    if not parts:
        from .capi import HIT_DTYPE
        return np.zeros(0, dtype=HIT_DTYPE)
    allh = np.concatenate(parts)
    return allh[np.lexsort((allh["col"], allh["pos"]))]
