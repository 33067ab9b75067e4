"""CPU tests of the tensor-core filter's recall guarantee (no GPU: b200scan_debug_fold runs the host-side folding of
blamm_b200/csrc/b200scan.cu -- fold_col / fold_col_z / fold_col_i8 -- on one column).

The filter may only ever ADD candidates: for every window whose in-order FP32 score (the reference's naive path,
motif.cpp:225-239) reaches the threshold (`hit <=> !(score < thr)`, pwmscan.cpp:115-117) the accumulator of the tensor pass must
come out >= 0.  Checked here on windows that matter: every window of short columns, the best window, random windows, and
boundary windows (random walks that degrade the best window while it stays a hit).

  * INT8 operands (the default): S32 accumulation is exact, so the integer sum IS the accumulator.
  * FP16 operands: the real sum of the FP16 weights must exceed the column's error allowance (margin - e1), and the accumulation
    itself is emulated under the two models the bound is derived for -- FP32 accumulators (in-order float32 adds) and FP16
    accumulators rounded DOWN after every single add (the worst case of "one FP16 ulp per internal add").
"""
import ctypes

import numpy as np
import pytest

from blamm_b200 import capi, synth

_dp = ctypes.POINTER(ctypes.c_double)


def debug_fold(w, thr, kind, zmode=0, margin16_scale=1.0):
    L = capi.scan_lib()
    fn = L.b200scan_debug_fold
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_float, ctypes.c_int32, ctypes.c_int32, ctypes.c_double,
                   ctypes.c_void_p, _dp, _dp, ctypes.POINTER(ctypes.c_int32)]
    w = np.ascontiguousarray(w, dtype=np.float32)
    out = np.zeros(w.shape, dtype=np.float64)
    bias, margin, flags = ctypes.c_double(), ctypes.c_double(), ctypes.c_int32()
    rc = fn(w.ctypes.data, w.shape[0], float(thr), kind, zmode, margin16_scale, out.ctypes.data, ctypes.byref(bias), ctypes.byref(margin),
            ctypes.byref(flags))
    assert rc == 0
    return out, bias.value, margin.value, flags.value


def inorder_scores(w, codes, mask=None):
    """float32 sum in position order; a masked position adds nothing (the reference's lower-case rule, sequence.cpp:312-319)."""
    s = np.zeros(len(codes), dtype=np.float32)
    for j in range(w.shape[0]):
        x = w[j, codes[:, j]]
        if mask is not None:
            x = np.where(mask[:, j], np.float32(0), x)
        s = (s + x).astype(np.float32)
    return s


def windows_for(w, thr, rng, n_random=20000, n_walks=300):
    L = w.shape[0]
    if L <= 8:
        idx = np.arange(4 ** L, dtype=np.int64)
        return np.stack([(idx >> (2 * j)) & 3 for j in range(L)], axis=1)
    best = np.argmax(w, axis=1)
    sets = [best[None, :], rng.integers(0, 4, size=(n_random, L))]
    # the best window with a few substitutions: mostly hits
    for k in (1, 2, 3, 4):
        m = np.repeat(best[None, :], 2000, axis=0)
        for _ in range(k):
            m[np.arange(len(m)), rng.integers(0, L, len(m))] = rng.integers(0, 4, len(m))
        sets.append(m)
    # boundary walks: degrade the best window at random positions while the in-order FP32 score stays >= thr
    cur = np.repeat(best[None, :], n_walks, axis=0)
    for step in range(4 * L):
        trial = cur.copy()
        trial[np.arange(n_walks), rng.integers(0, L, n_walks)] = rng.integers(0, 4, n_walks)
        keep = ~(inorder_scores(w, trial) < np.float32(thr))
        cur[keep] = trial[keep]
        if step % 2:
            sets.append(cur.copy())
    return np.concatenate(sets)


def _columns(tmp_path, n=20, seed=77):
    path = str(tmp_path / "m.jaspar")
    synth.make_jaspar_like(path, n, seed)
    ms = capi.MotifSet(path, revcompl=True)
    P, col_len, _ = ms.generate_matrix([2950, 2050, 2050, 2950])
    mn, mx = ms.min_max()
    return [(P[c, :4 * col_len[c]].reshape(col_len[c], 4).copy(), float(mn[c]), float(mx[c])) for c in range(ms.n_cols)]


def _thresholds(w, mn, mx, rng):
    """relative thresholds, a p-value-like one (the 1e-4 quantile of random windows), and the edges of the attainable range"""
    L = w.shape[0]
    rnd = inorder_scores(w, rng.integers(0, 4, size=(30000, L)))
    best = inorder_scores(w, np.argmax(w, axis=1)[None, :])[0]
    return [np.float32(0.95 * (mx - mn) + mn), np.float32(0.8 * (mx - mn) + mn), np.float32(np.quantile(rnd, 1 - 1e-4)),
            best, np.nextafter(best, np.float32(-np.inf)), np.nextafter(best, np.float32(np.inf)), np.float32(0.0)]


def test_int8_filter_never_loses_a_hit(tmp_path):
    rng = np.random.default_rng(1)
    checked = hits_seen = unresolved = 0
    for w, mn, mx in _columns(tmp_path):
        for thr in _thresholds(w, mn, mx, rng):
            q, _, overshoot, flags = debug_fold(w, thr, 8)
            codes = windows_for(w, thr, rng)
            hit = ~(inorder_scores(w, codes) < thr)
            if flags & 2:                                   # "no window can reach the threshold"
                assert not hit.any()
                continue
            if flags & 1:                                   # INT8 cannot resolve it: every window is a candidate anyway
                unresolved += 1
                continue
            assert np.all(q == np.round(q)) and np.abs(q).max() <= 127
            acc = q[np.arange(w.shape[0])[None, :], codes].sum(axis=1)
            assert np.all(acc[hit] >= 0), (w.shape[0], float(thr), float(acc[hit].min()))
            assert np.abs(acc).max() < 2 ** 15              # what lets the epilogue read S32 accumulators as packed S16
            # the filter is not vacuous: a window that passes although its score lies more than the worst-case rounding overshoot
            # (L / scale) below the threshold owes that to a letter clamped at -127 (a letter that loses more than the whole slack)
            picked = q[np.arange(w.shape[0])[None, :], codes]
            far_below = inorder_scores(w, codes) <= thr - np.float32(overshoot) - np.float32(1e-2)
            assert overshoot > 0 and np.all(~far_below | (acc < 0) | (picked == -127).any(axis=1))
            checked += 1
            hits_seen += int(hit.sum())
    assert checked > 150 and hits_seen > 100000 and unresolved < checked


def test_int8_filter_with_masked_positions(tmp_path):
    """blocks with zero-contribution characters: unshifted integer weights + the bias step; a masked position adds exactly 0"""
    rng = np.random.default_rng(2)
    checked = hits_seen = 0
    for w, mn, mx in _columns(tmp_path, n=14, seed=78):
        for thr in _thresholds(w, mn, mx, rng):
            q, bias, _, flags = debug_fold(w, thr, 8, zmode=1)
            codes = windows_for(w, thr, rng, n_random=8000, n_walks=150)
            if w.shape[0] <= 8:
                codes = codes[rng.integers(0, len(codes), 30000)]
            mask = rng.random(codes.shape) < rng.choice([0.0, 0.1, 0.5], size=(len(codes), 1))
            hit = ~(inorder_scores(w, codes, mask) < thr)
            if flags & 2:
                assert not hit.any()
                continue
            if flags & 1:
                continue
            assert bias == round(bias) and -4064 <= bias <= 4064
            acc = bias + np.where(mask, 0.0, q[np.arange(w.shape[0])[None, :], codes]).sum(axis=1)
            assert np.all(acc[hit] >= 0), (w.shape[0], float(thr), float(acc[hit].min()))
            # a fully masked window scores 0: a candidate only if 0 reaches the threshold
            if thr > 1e-2:
                assert bias < 0
            checked += 1
            hits_seen += int(hit.sum())
    assert checked > 40 and hits_seen > 30000


def _round_down_f16(x):
    h = x.astype(np.float16)
    over = h.astype(np.float64) > x
    return np.where(over, np.nextafter(h, np.float16(-np.inf)), h).astype(np.float64)


@pytest.mark.parametrize("zmode", [0, 1])
def test_fp16_operand_filter_never_loses_a_hit(tmp_path, zmode):
    rng = np.random.default_rng(3 + zmode)
    checked = hits_seen = 0
    for w, mn, mx in _columns(tmp_path, n=10, seed=79):
        L = w.shape[0]
        A = float(np.abs(w).max(axis=1).astype(np.float64).sum())
        e1 = (L - 1) * A * 2.0 ** -24
        for thr in _thresholds(w, mn, mx, rng):
            codes = windows_for(w, thr, rng, n_random=6000, n_walks=120)
            if L <= 8:
                codes = codes[rng.integers(0, len(codes), 20000)]
            mask = (rng.random(codes.shape) < rng.choice([0.0, 0.1, 0.5], size=(len(codes), 1))) if zmode else None
            hit = ~(inorder_scores(w, codes, mask) < thr)
            for kind in (32, 16):
                y, bias, margin, flags = debug_fold(w, thr, kind, zmode)
                if flags & 1 or margin > 1e8:               # degenerate column / bound did not converge: handled as "always" or FP32 by the caller
                    continue
                picked = y[np.arange(L)[None, :], codes]
                if mask is not None:
                    picked = np.where(mask, 0.0, picked)
                real = bias + picked.sum(axis=1)
                assert np.all(real[hit] >= margin - e1 - 1e-9), (L, float(thr), kind)
                if kind == 32:                              # FP32 accumulators: in-order float32 adds
                    acc = np.full(len(codes), bias, dtype=np.float32)
                    for j in range(L):
                        acc = (acc + picked[:, j].astype(np.float32)).astype(np.float32)
                else:                                       # FP16 accumulators, rounded DOWN after every add
                    acc = np.full(len(codes), bias, dtype=np.float64)
                    for j in range(L):
                        acc = _round_down_f16(acc + picked[:, j])
                assert np.all(acc[hit] >= 0), (L, float(thr), kind, float(np.min(acc[hit])))
                checked += 1
            hits_seen += int(hit.sum())
    assert checked > 50 and hits_seen > 20000


def test_debug_fold_rejects_bad_arguments():
    L = capi.scan_lib()
    w = np.zeros((3, 4), dtype=np.float32)
    out = np.zeros((3, 4)); d = ctypes.c_double(); f = ctypes.c_int32()
    fn = L.b200scan_debug_fold
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_float, ctypes.c_int32, ctypes.c_int32, ctypes.c_double,
                   ctypes.c_void_p, _dp, _dp, ctypes.POINTER(ctypes.c_int32)]
    assert fn(w.ctypes.data, 0, 1.0, 8, 0, 1.0, out.ctypes.data, ctypes.byref(d), ctypes.byref(d), ctypes.byref(f)) == -1
    assert fn(w.ctypes.data, 65, 1.0, 8, 0, 1.0, out.ctypes.data, ctypes.byref(d), ctypes.byref(d), ctypes.byref(f)) == -1
    assert fn(w.ctypes.data, 3, 1.0, 7, 0, 1.0, out.ctypes.data, ctypes.byref(d), ctypes.byref(d), ctypes.byref(f)) == -1
    assert fn(None, 3, 1.0, 8, 0, 1.0, out.ctypes.data, ctypes.byref(d), ctypes.byref(d), ctypes.byref(f)) == -1
