#!/usr/bin/env python
"""BASELINE.json configs[3] at reduced length: 10,000 synthetic PWMs of length 6-30 (20,000 columns with -rc), absolute
threshold 12, uniform ACGT.  Checks that tensor and gather engines agree and reports throughput.  usage: config4.py [Mbp]"""
import os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blamm_b200 import capi, synth

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
n = int(mbp * 1e6)
work = tempfile.mkdtemp()
mf = os.path.join(work, "m.jaspar")
t0 = time.time(); synth.make_jaspar_like(mf, 10000, 77, uniform_len=(6, 30)); print("motif file: %.1f s" % (time.time() - t0))
seq = synth.random_acgt(n, 99)
ms = capi.MotifSet(mf, revcompl=True)
P, col_len, is_rc = ms.generate_matrix(synth.counts_of(seq[:2_000_000]))
thr = ms.thresholds("at", 12.0)
print("columns", len(col_len), "sum L", int(col_len.sum()), "mean L %.2f" % col_len.mean())
sc = capi.Scanner(0, max_block_nt=n + 64, max_hits=1 << 24)
res = {}
for name, eng in (("tensor", capi.ENGINE_TENSOR), ("gather", capi.ENGINE_GATHER)):
    sc.set_engine(eng)
    t0 = time.time(); sc.set_motifs(P, col_len, thr); t_set = time.time() - t0
    hits, t = sc.scan(seq)
    hits, t = sc.scan(seq)
    res[name] = hits[np.lexsort((hits["col"], hits["pos"]))]
    print("%s: %d hits, cand %d, score %.2f ms + rescore %.2f ms -> %.3e scores/s (set_motifs %.2f s, tiles %d, %s)" % (
        name, len(hits), t["n_candidates"], t["score_ms"], t["rescore_ms"], n * len(col_len) / ((t["score_ms"] + t["rescore_ms"]) * 1e-3), t_set,
        sc.describe()["n_tiles"], sc.tensor_info() if eng == capi.ENGINE_TENSOR else ""))
print("engines identical:", np.array_equal(res["tensor"], res["gather"]))
