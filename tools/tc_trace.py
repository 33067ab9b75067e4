#!/usr/bin/env python
"""Per-tile timeline of the filter kernel's warp roles from the B200_TRACE build (clock64 probes, CTA 0, item 0).
usage: B200SCAN_LIB=blamm_b200/lib/libb200scan_trace.so python tools/tc_trace.py [n_motifs] [Lmin] [Lmax] [acc]"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200SCAN_LIB", os.path.join(ROOT, "blamm_b200", "lib", "libb200scan_trace.so"))
from blamm_b200 import capi
from tests import util

nm, lo, hi, acc = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 128), (2, 10), (3, 12), (4, 0)))
case = util.random_case(5, n_motifs=nm, n_nt=8_000_000, len_range=(lo, hi), with_gaps=False)
case["thr"] = np.maximum(case["thr"], 9.0).astype(np.float32)
sc = capi.Scanner(0, max_block_nt=len(case["chars"]) + 64, max_hits=1 << 22)
sc.set_engine(capi.ENGINE_TENSOR); sc.set_tensor_accumulator(acc)
sc.set_motifs(case["P"], case["col_len"], case["thr"])
hits, t = sc.scan(case["chars"])
print("cols", len(case["col_len"]), "tiles", sc.describe()["n_tiles"], sc.tensor_info(), "hits", len(hits), "cand", t["n_candidates"], "score_ms %.3f" % t["score_ms"])
n = 4 * 256 * 4
buf = (ctypes.c_uint64 * n)()
L = capi.scan_lib(); L.b200scan_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
assert L.b200scan_debug_trace(sc._ctx, buf, n) == 0
T = np.frombuffer(buf, dtype=np.uint64).reshape(4, 256, 4).astype(np.int64)
t0 = T[T > 0].min()
T = np.where(T > 0, T - t0, -1)
print("tile |  producer(wait,done) |  MMA(eFull, tEmpty, issued) | epi first warp (tFull, -, released, done) | epi 8th warp (...)   [with 2 epilogue groups a warp only sees every other tile]")
for i in list(range(0, 12)) + list(range(100, 112)):
    print("%4d | %6d %6d | %6d %6d %6d | %6d %6d %6d %6d | %6d %6d %6d %6d" % ((i,) + tuple(T[0, i, :2]) + tuple(T[1, i, :3]) + tuple(T[2, i]) + tuple(T[3, i])))
ev = np.arange(60, 200, 2)          # tiles of group 0 (even) -- valid for one or two epilogue groups
od = ev + 1
per = np.diff(T[1, 60:200, 2]).mean()
def m(x): return float(np.mean(x))
print("steady state: period %.0f cyc/tile | MMA warp: eFull->tEmpty wait %.0f, tEmpty->issued %.0f | producer: %.0f busy, %.0f waiting" % (
    per, m(T[1, 60:200, 1] - T[1, 60:200, 0]), m(T[1, 60:200, 2] - T[1, 60:200, 1]), m(T[0, 60:200, 1] - T[0, 60:200, 0]), m(T[0, 61:201, 0] - T[0, 60:200, 1])))
for name, r, tiles in (("first epilogue warp", 2, ev), ("8th epilogue warp", 3, od if T[3, 61, 0] > 0 else ev)):
    print("  %s: MMA issued -> tFull seen %.0f | tFull -> released %.0f | released -> done %.0f | done -> its next tFull %.0f" % (
        name, m(T[r, tiles, 0] - T[1, tiles, 2]), m(T[r, tiles, 2] - T[r, tiles, 0]), m(T[r, tiles, 3] - T[r, tiles, 2]),
        m(T[r, tiles[1:], 0] - T[r, tiles[:-1], 3])))
    print("     released(i) -> MMA sees tEmpty for tile i+2: %.0f" % m(T[1, tiles[:-1] + 2, 1] - T[r, tiles[:-1], 2]))
