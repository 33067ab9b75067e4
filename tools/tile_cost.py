#!/usr/bin/env python
"""Cost of one 128-window tile of the tensor filter as a function of its width N and of its MMA steps (the inputs of
b200scan.cu: plan_tc_tiles).  One column tile per run: n_cols columns of equal length.  usage: tile_cost.py [Mbp]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blamm_b200 import capi, synth

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 50.0
n = int(mbp * 1e6)
seq = synth.random_acgt(n, 5)
sc = capi.Scanner(0, max_block_nt=n + 64, max_hits=1 << 24)
rng = np.random.default_rng(1)
for L in (8, 16, 32):
    for n_cols in (64, 128, 192, 256):
        P = np.zeros((n_cols, 4 * L), dtype=np.float32)
        W = rng.uniform(-3.0, 0.0, size=(n_cols, L, 4)).astype(np.float32)
        best = rng.integers(0, 4, size=(n_cols, L))
        np.put_along_axis(W, best[:, :, None], rng.uniform(0.5, 1.5, size=(n_cols, L, 1)).astype(np.float32), axis=2)
        P[:] = W.reshape(n_cols, -1)
        thr = (W.max(axis=2).sum(axis=1) * 0.8).astype(np.float32)
        col_len = np.full(n_cols, L, dtype=np.int32)
        for bits in (8, 16):
            sc.set_engine(capi.ENGINE_TENSOR); sc.set_tensor_accumulator(bits)
            sc.set_motifs(P, col_len, thr)
            hits, t = sc.scan(seq)
            tot, k_ms, nh = sc.rerun_resident(0, 3)
            tiles_per_sm = (n / 128.0) * sc.describe()["n_tiles"] / sc.describe()["sm_count"]
            print("L %2d  N %3d  acc %2d: tiles %d  %.3f ms per launch -> %4.0f cycles per tile at 1.9 GHz  (%d hits, %d candidates)" % (
                L, n_cols, bits, sc.describe()["n_tiles"], k_ms / 3, k_ms / 3 * 1e-3 * 1.9e9 / tiles_per_sm, len(hits), t["n_candidates"]))
sc.close()
