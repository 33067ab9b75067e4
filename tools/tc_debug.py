#!/usr/bin/env python
"""GPU-side diagnostic: tensor-core filter vs gather-add on seeded cases (hit-set diff by column length) and
a quick throughput probe of both engines.  Not part of the product or the test-suite."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blamm_b200 import capi  # noqa: E402
from tests import util  # noqa: E402


def diff(case, n_nt_label, acc=0):
    sc = capi.Scanner(0, max_block_nt=max(1 << 20, len(case["chars"]) + 64), max_hits=1 << 22)
    out = {}
    for name, eng in (("gather", capi.ENGINE_GATHER), ("tensor", capi.ENGINE_TENSOR)):
        sc.set_engine(eng)
        sc.set_tensor_accumulator(acc)
        sc.set_motifs(case["P"], case["col_len"], case["thr"])
        if name == "tensor":
            print("  tensor info:", sc.tensor_info())
        t0 = time.time()
        try:
            hits, t = sc.scan(case["chars"], case["frag_start"][1:])
        except Exception as e:
            print("  %s: FAILED %s" % (name, e))
            return
        hits = hits[np.lexsort((hits["col"], hits["pos"]))]
        out[name] = hits
        print("  %s %s: %d hits, cand %d, score %.3f ms rescore %.3f ms, wall %.3f s, tiles %s" % (
            n_nt_label, name, len(hits), t["n_candidates"], t["score_ms"], t["rescore_ms"], time.time() - t0, sc.describe()))
    g, tt = out["gather"], out["tensor"]
    if np.array_equal(g, tt):
        print("  IDENTICAL")
    else:
        gs = set(zip(g["pos"].tolist(), g["col"].tolist())); ts = set(zip(tt["pos"].tolist(), tt["col"].tolist()))
        miss, extra = sorted(gs - ts), sorted(ts - gs)
        print("  MISMATCH: missing %d extra %d" % (len(miss), len(extra)))
        L = case["col_len"]
        import collections
        print("   missing by len:", sorted(collections.Counter(int(L[c]) for _, c in miss).items()))
        print("   missing pos%128 sample:", [p % 128 for p, _ in miss[:20]], " first:", miss[:8])
        print("   extra sample:", extra[:8])
    sc.close()


if __name__ == "__main__":
    print("case A: 12 motifs L 5..12, 50k nt")
    diff(util.random_case(1, n_motifs=12, n_nt=50_000, len_range=(5, 12), with_gaps=False), "50k")
    print("case B: 30 motifs L 5..30, 300k nt, gaps")
    diff(util.random_case(2, n_motifs=30, n_nt=300_000), "300k")
    print("case C: 300 motifs L 5..35, 8M nt")
    c = util.random_case(3, n_motifs=300, n_nt=8_000_000, len_range=(5, 35))
    c["thr"] = np.maximum(c["thr"], 10.0).astype(np.float32)
    for acc in (32, 16):
        diff(c, "8M acc%d" % acc, acc)
    # how much of the FP16-accumulation error bound does the hardware actually use?  Shrink the bound and count misses.
    for scale in ("0.25", "0.05", "0.0"):
        os.environ["B200SCAN_MARGIN16_SCALE"] = scale
        print("FP16 accumulators with the error bound scaled by", scale)
        diff(c, "8M acc16 x" + scale, 16)
    os.environ["B200SCAN_MARGIN16_SCALE"] = "1.0"
