"""Parity of the CUDA path (through the C ABI, libb200scan.so) with the oracle and with the golden fixtures
produced by the compiled reference.  Bar: identical occurrence set; scores bit-identical to the oracle's
in-order FP32 sum (== the reference's naive path) and within 1e-5 of the reference's BLAS path."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from blamm_b200 import capi, lib_dir, shard, synth
from oracle import oracle as O
from oracle.refdump_io import read_refdump
from tests import util

pytestmark = pytest.mark.gpu
ENGINES = [capi.ENGINE_GATHER, capi.ENGINE_TENSOR, capi.ENGINE_AUTO]


@pytest.fixture(scope="module")
def scanner():
    s = capi.Scanner(0, max_block_nt=1 << 25, max_hits=1 << 20)
    yield s
    s.close()


def _sorted(h):
    return h[np.lexsort((h["col"], h["pos"]))]


def _oracle_hits(case, n_payload=None, lower_fold=False):
    pos, col, sc = O.scan_stream(bytes(case["chars"]), case["frag_start"], case["P"], case["col_len"], case["thr"],
                                 n_payload=n_payload, lower_fold=lower_fold)
    return pos, col, sc


def _assert_same(hits, pos, col, sc):
    h = _sorted(hits)
    assert len(h) == len(pos), "hit count %d vs oracle %d" % (len(h), len(pos))
    assert np.array_equal(h["pos"], pos) and np.array_equal(h["col"], col)
    assert np.array_equal(h["score"].view(np.uint32), sc.view(np.uint32))       # bit-exact (integer compare of the FP32 bits)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("case_name,mode_key", [("example", "pt_rc"), ("example", "pt_fwd"), ("example", "rt_rc"), ("example", "at_rc"),
                                                 ("edge", "pt_rc"), ("edge", "rt_rc"), ("edge", "at_low")])
def test_golden_cases(scanner, golden, engine, case_name, mode_key):
    """Full `blamm scan` semantics on the fixtures: host model (C++) -> C ABI -> hits -> occurrence lines."""
    d = os.path.join(golden, case_name)
    mode, value, rc = util.MODES[mode_key]
    ms = capi.MotifSet(os.path.join(d, "motifs.jaspar"), rc)
    dump = read_refdump(os.path.join(d, "refdump_%s.bin" % mode_key))
    has_lower = case_name == "edge"
    lines = []
    scanner.set_engine(engine)
    for sp, r in zip(O.load_dict(os.path.join(d, "sequences.mf.dict")), dump):
        P, col_len, is_rc = ms.generate_matrix(sp.counts)
        thr = ms.thresholds(mode, value, sp.name, d)
        scanner.set_motifs(P, col_len, thr)
        fs = capi.FastaStream([os.path.join(d, f) for f in sp.files], sp.tot_len)
        halo = int(col_len.max()) - 1
        got = []
        while True:
            c = fs.next(20000, halo)            # several chunks per group: exercises the halo hand-over
            if c is None:
                break
            hits, t = scanner.scan(c["chars"], c["frag_start"][1:], c["n_payload"])
            f = np.searchsorted(c["frag_start"], hits["pos"], side="right") - 1
            for h, fi in zip(hits, f):
                got.append((int(c["frag_seq"][fi]), int(c["frag_pos"][fi] + h["pos"] - c["frag_start"][fi]), int(h["col"]),
                            int(np.float32(h["score"]).view(np.uint32))))
        ref = r["hits"]
        assert sorted(g[:3] for g in got) == sorted(zip(ref["seq"].tolist(), ref["pos"].tolist(), ref["col"].tolist()))
        got.sort()
        refs = ref[np.lexsort((ref["col"], ref["pos"], ref["seq"]))]
        mine = np.array([g[3] for g in got], dtype=np.uint32).view(np.float32)
        assert np.max(np.abs(mine - refs["score"]), initial=0.0) <= 1e-5          # vs the reference's BLAS path
        for g in got:
            nm, L = ms.names[g[2]], int(col_len[g[2]])
            lines.append("%s\tblamm\t%s\t%d\t%d\t%s\t%s\t.\t.\n" % (sp.seq_names[g[0]], nm, g[1], g[1] + L,
                                                                 O.fmt_g(np.uint32(g[3]).view(np.float32)), "-" if is_rc[g[2]] else "+"))
    if case_name == "example":      # byte-identical to the reference's occurrences.txt (sorted)
        assert sorted(lines) == open(os.path.join(d, "occ_%s.txt" % mode_key)).read().splitlines(True)


@pytest.mark.parametrize("engine,acc", [(capi.ENGINE_GATHER, 0), (capi.ENGINE_TENSOR, 8), (capi.ENGINE_TENSOR, 16), (capi.ENGINE_TENSOR, 32)])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_cases_bit_exact(scanner, engine, acc, seed):
    case = util.random_case(seed, n_motifs=30, n_nt=300_000)
    scanner.set_engine(engine)
    scanner.set_tensor_accumulator(acc)
    scanner.set_motifs(case["P"], case["col_len"], case["thr"])
    if acc:
        assert scanner.tensor_info()["accumulator_bits"] == acc
    hits, t = scanner.scan(case["chars"], case["frag_start"][1:])
    scanner.set_tensor_accumulator(0)
    _assert_same(hits, *_oracle_hits(case))
    assert t["engine_used"] == engine and t["kernel_launches"] >= 2
    if acc:      # the filter is conservative (no hit may be lost) and tight (few wasted candidates)
        assert len(hits) <= t["n_candidates"] <= (2.5 if acc != 8 else 5.0) * len(hits) + 100      # INT8 weights: < 1 / scale overshoot per position


@pytest.mark.parametrize("lower", [capi.LOWER_ZERO, capi.LOWER_FOLD])
def test_lower_case_semantics(scanner, lower):
    case = util.random_case(11, n_motifs=16, n_nt=120_000, lower=True)
    scanner.set_engine(capi.ENGINE_AUTO)
    scanner.set_motifs(case["P"], case["col_len"], case["thr"])
    hits, t = scanner.scan(case["chars"], case["frag_start"][1:], lower=lower)
    _assert_same(hits, *_oracle_hits(case, lower_fold=(lower == capi.LOWER_FOLD)))
    assert t["engine_used"] == capi.ENGINE_TENSOR        # masked blocks have their own tensor instance (bias step, zeroed rows)


@pytest.mark.parametrize("engine,acc", [(capi.ENGINE_GATHER, 0), (capi.ENGINE_TENSOR, 8), (capi.ENGINE_TENSOR, 16), (capi.ENGINE_TENSOR, 32)])
@pytest.mark.parametrize("seed", [5, 6])
def test_soft_masked_blocks_bit_exact(scanner, engine, acc, seed):
    """Soft-masked sequence (half of it lower case, in runs of 1 .. 3000, as repeat-masked genomes are): lower-case
    characters contribute exactly 0 (the reference's BLAS path, sequence.cpp:312-319), negative and positive thresholds."""
    case = util.random_case(seed, n_motifs=40, n_nt=700_000, len_range=(5, 40) if seed == 5 else (5, 64))      # 64: the bias step on top of 16 position steps
    rng = np.random.default_rng(seed + 100)
    chars = case["chars"].copy()
    p = 0
    while p < len(chars):
        run = int(rng.integers(1, 3000))
        if rng.random() < 0.5:
            chars[p:p + run] |= 0x20                     # lower case
        p += run
    case["chars"] = chars
    thr = case["thr"].copy()
    thr[::8] = np.float32(-1.5)                          # every 8th column: negative threshold (fully masked windows score 0 >= thr)
    thr[1::3] = np.minimum(thr[1::3], 4.0)
    case["thr"] = thr
    scanner.set_engine(engine)
    scanner.set_tensor_accumulator(acc)
    scanner.set_motifs(case["P"], case["col_len"], case["thr"])
    hits, t = scanner.scan(case["chars"], case["frag_start"][1:])
    scanner.set_tensor_accumulator(0)
    want = _oracle_hits(case)
    assert len(want[0]) > 10_000
    _assert_same(hits, *want)
    assert t["engine_used"] == engine


@pytest.mark.parametrize("engine", [capi.ENGINE_GATHER, capi.ENGINE_TENSOR])
def test_edge_blocks(scanner, engine):
    case = util.random_case(21, n_motifs=12, n_nt=5000, len_range=(6, 40))
    scanner.set_engine(engine)
    scanner.set_motifs(case["P"], case["col_len"], case["thr"])
    maxlen = int(case["col_len"].max())
    # empty block, block shorter than every motif, block of exactly one window, ragged payload/halo splits
    hits, _ = scanner.scan(b"", None)
    assert len(hits) == 0
    for n in (1, 3, maxlen - 1, maxlen, maxlen + 1, 127, 128, 129, 4097):
        sub = dict(case, chars=case["chars"][:n], frag_start=case["frag_start"][case["frag_start"] < n])
        hits, _ = scanner.scan(sub["chars"], sub["frag_start"][1:])
        _assert_same(hits, *_oracle_hits(sub))
    for n_payload in (0, 1, 100, 4999 - maxlen, 5000):
        hits, _ = scanner.scan(case["chars"], case["frag_start"][1:], n_payload)
        _assert_same(hits, *_oracle_hits(case, n_payload=n_payload))


@pytest.mark.parametrize("engine", [capi.ENGINE_GATHER, capi.ENGINE_TENSOR])
def test_max_length_and_many_columns(scanner, engine):
    """Motifs up to the ABI limit of 64 positions and >256 columns (several tensor tiles, several gather tiles)."""
    case = util.random_case(31, n_motifs=150, n_nt=60_000, len_range=(4, 64))
    assert case["col_len"].max() == 64 and len(case["col_len"]) == 300
    for acc in ((16, 32) if engine == capi.ENGINE_TENSOR else (0,)):
        scanner.set_engine(engine)
        scanner.set_tensor_accumulator(acc)
        scanner.set_motifs(case["P"], case["col_len"], case["thr"])
        hits, _ = scanner.scan(case["chars"], case["frag_start"][1:])
        _assert_same(hits, *_oracle_hits(case))
    scanner.set_tensor_accumulator(0)


def test_accumulator_type_is_chosen_per_tile(scanner):
    """Short motifs run on INT8 operands (or, with those excluded, FP16 accumulators) while a tile of long motifs with extreme
    weights and a very low threshold (INT8 overshoot and FP16 error bound too large) falls back to FP16 operands with FP32
    accumulators: the set is 'mixed' (tensor_info reports 0) and the hit list still equals the oracle's."""
    case = util.random_case(91, n_motifs=150, n_nt=300_000, len_range=(6, 12))
    rng = np.random.default_rng(92)
    n_short, ldp = case["P"].shape[0], 4 * 64
    P = np.zeros((n_short + 8, ldp), dtype=np.float32)
    P[:n_short, :case["P"].shape[1]] = case["P"]
    col_len = np.concatenate([case["col_len"], np.full(8, 64, dtype=case["col_len"].dtype)])
    thr = np.concatenate([case["thr"], np.zeros(8, dtype=np.float32)])
    codes = np.frombuffer(bytes(case["chars"][:4096]), dtype=np.uint8)
    lut = np.zeros(256, dtype=np.int64); lut[ord("C")] = 1; lut[ord("G")] = 2; lut[ord("T")] = 3
    for k in range(8):                                   # weights in [-20, 2]; the window at 100 + 37 k scores the maximum
        W = rng.uniform(-20.0, -1.0, size=(64, 4)).astype(np.float32)
        best = lut[codes[100 + 37 * k: 164 + 37 * k]]
        W[np.arange(64), best] = rng.uniform(0.5, 2.0, size=64).astype(np.float32)
        P[n_short + k, :] = W.reshape(-1)
        thr[n_short + k] = np.float32(-400.0)            # ~2 sigma above the mean window score: thousands of hits per column
    scanner.set_engine(capi.ENGINE_TENSOR)
    scanner.set_tensor_accumulator(0)
    scanner.set_motifs(P, col_len, thr)
    assert scanner.tensor_info()["accumulator_bits"] == 0, scanner.tensor_info()
    hits, t = scanner.scan(case["chars"], case["frag_start"][1:])
    want = _oracle_hits(dict(case, P=P, col_len=col_len, thr=thr))
    assert (hits["col"] >= n_short).sum() > 100
    _assert_same(hits, *want)
    for bits in (8, 16, 32):                             # forcing any type everywhere changes nothing
        scanner.set_tensor_accumulator(bits)
        scanner.set_motifs(P, col_len, thr)
        assert scanner.tensor_info()["accumulator_bits"] == bits
        h2, _ = scanner.scan(case["chars"], case["frag_start"][1:])
        _assert_same(h2, *want)
    scanner.set_tensor_accumulator(0)


@pytest.mark.parametrize("lower", [False, True])
def test_int8_filter_at_the_edges_of_its_bound(scanner, lower):
    """The INT8 instance rounds and clamps its integer weights UP, so it must never lose a hit -- also where its
    quantisation is at its limits: thresholds exactly at, one ulp below and above the best attainable score (slack 0:
    scale at its maximum, every other letter clamps), thresholds just below that (a single mild mismatch allowed),
    extreme weights (-60 ... +8: clamped letters), thresholds far below the best score (INT8 cannot resolve them: forced
    INT8 sends every window to the exact rescorer), negative thresholds, and lengths 1, 8, 9, 64 around the 8-position MMA step."""
    rng = np.random.default_rng(321)
    n_nt = 400_000
    chars = synth.random_acgt(n_nt, 77)
    lut = np.zeros(256, dtype=np.int64); lut[ord("C")] = 1; lut[ord("G")] = 2; lut[ord("T")] = 3
    codes = lut[chars]
    lens = [1, 2, 7, 8, 9, 15, 16, 17, 24, 25, 33, 64] * 4
    ldp = 4 * 64
    P = np.zeros((len(lens), ldp), dtype=np.float32)
    thr = np.zeros(len(lens), dtype=np.float32)
    for c, L in enumerate(lens):
        kind = c // 12
        W = rng.uniform(-6.0, 1.9, size=(L, 4)).astype(np.float32)
        if kind == 1:
            W = rng.uniform(-60.0, 8.0, size=(L, 4)).astype(np.float32)
        at = int(rng.integers(0, n_nt - 64))                     # plant the column's best window in the sequence
        W[np.arange(L), codes[at:at + L]] = rng.uniform(0.4, 2.0, size=L).astype(np.float32) * (4.0 if kind == 1 else 1.0)
        P[c, :4 * L] = W.reshape(-1)
        best = np.float32(0.0)
        for j in range(L):                                       # the reference's in-order FP32 sum of the best letters
            best = np.float32(best + W[j].max())
        if kind == 0:
            thr[c] = [best, np.nextafter(best, np.float32(-np.inf)), np.nextafter(best, np.float32(np.inf))][c % 3]
        elif kind == 1:
            thr[c] = np.float32(best - rng.uniform(0.0, 3.0))
        elif kind == 2:
            thr[c] = np.float32(best - (0.5 + 0.25 * L))
        else:
            thr[c] = np.float32(-1.0 - 0.1 * L) if L > 2 else np.float32(0.3)
    col_len = np.array(lens, dtype=np.int32)
    if lower:
        chars = chars.copy()
        chars[1000:200_000:3] |= 0x20
    case = dict(P=P, col_len=col_len, thr=thr, chars=chars, frag_start=np.array([0, 5, 300_001], dtype=np.uint64))
    want = _oracle_hits(case)
    assert len(want[0]) > 1000
    big = capi.Scanner(0, max_block_nt=1 << 20, max_hits=1 << 23)
    try:
        for bits in (8, 0):
            big.set_engine(capi.ENGINE_TENSOR)
            big.set_tensor_accumulator(bits)
            big.set_motifs(P, col_len, thr)
            hits, t = big.scan(chars, case["frag_start"][1:])
            _assert_same(hits, *want)
    finally:
        big.close()


@pytest.mark.parametrize("engine", [capi.ENGINE_GATHER, capi.ENGINE_TENSOR])
def test_many_short_fragments(scanner, engine):
    """120,000 sequences of 1 .. 40 characters in one block (most shorter than the motifs): the fragment rule of
    SeqBlock::getRemainingSeqLen decides almost every window."""
    case = util.random_case(33, n_motifs=30, n_nt=2_400_000, len_range=(5, 24), with_gaps=False)
    rng = np.random.default_rng(34)
    starts = np.cumsum(rng.integers(1, 41, size=120_000)).astype(np.uint64)
    case["frag_start"] = np.concatenate([np.zeros(1, dtype=np.uint64), starts[starts < len(case["chars"]) - 1]])
    case["thr"] = np.minimum(case["thr"], 6.0).astype(np.float32)
    scanner.set_engine(engine)
    scanner.set_motifs(case["P"], case["col_len"], case["thr"])
    hits, t = scanner.scan(case["chars"], case["frag_start"][1:])
    want = _oracle_hits(case)
    assert len(want[0]) > 1000
    _assert_same(hits, *want)


def test_cta_pair_kernel_matches(monkeypatch):
    """The optional CTA-pair instance of the filter (B200SCAN_PAIR=1: cta_group::2 MMAs issued by one lane for two SMs,
    completion multicast to both CTAs, hand-backs through the cluster shared window) returns the same hit lists."""
    monkeypatch.setenv("B200SCAN_PAIR", "1")
    sc = capi.Scanner(0, max_block_nt=1 << 22, max_hits=1 << 22)
    try:
        for seed, lower in ((3, False), (12, True)):
            case = util.random_case(seed, n_motifs=120, n_nt=3_000_001, len_range=(5, 40), lower=lower)
            sc.set_engine(capi.ENGINE_TENSOR)
            want = _oracle_hits(case)
            for bits in (16, 0):                         # the pair instance exists for FP16 operands; INT8 tiles (auto) run single-CTA
                sc.set_tensor_accumulator(bits)
                sc.set_motifs(case["P"], case["col_len"], case["thr"])
                hits, t = sc.scan(case["chars"], case["frag_start"][1:])
                _assert_same(hits, *want)
    finally:
        sc.close()


def test_packed_submit_and_zero_mask(scanner):
    case = util.random_case(41, n_motifs=10, n_nt=50_000)
    chars = case["chars"]
    code = np.zeros(256, dtype=np.uint32); code[ord("C")] = 1; code[ord("G")] = 2; code[ord("T")] = 3
    c = code[chars]
    n = len(c)
    padded = np.zeros((n + 15) // 16 * 16, dtype=np.uint32); padded[:n] = c
    words = (padded.reshape(-1, 16) << (2 * np.arange(16, dtype=np.uint32))).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    scanner.set_engine(capi.ENGINE_AUTO)
    scanner.set_motifs(case["P"], case["col_len"], case["thr"])
    scanner.submit_packed(0, words, None, n, n, case["frag_start"][1:])
    hits, t = scanner.collect(0)
    _assert_same(hits, *_oracle_hits(case))
    # zero mask: positions 1000..1999 contribute nothing == the oracle's lower-case (BLAS path) semantics
    zm = np.zeros((n + 31) // 32, dtype=np.uint32)
    for p in range(1000, 2000):
        zm[p // 32] |= np.uint32(1 << (p % 32))
    scanner.submit_packed(1, words, zm, n, n, case["frag_start"][1:])
    hits, t = scanner.collect(1)
    low = chars.copy(); low[1000:2000] = np.frombuffer(bytes(low[1000:2000]).lower(), dtype=np.uint8)
    _assert_same(hits, *_oracle_hits(dict(case, chars=low)))
    assert t["engine_used"] == capi.ENGINE_TENSOR


@pytest.mark.parametrize("lower", [capi.LOWER_ZERO, capi.LOWER_FOLD])
def test_host_packed_block_matches_ascii_block(scanner, lower):
    """blamm_pack_ascii (host twin of pack_ascii_kernel) + b200scan_submit_packed == b200scan_submit_ascii == oracle, for a block
    with lower-case stretches under both lower-case rules and a length that fills neither a code nor a mask word."""
    case = util.random_case(43, n_motifs=12, n_nt=123_457, lower=True)
    chars = case["chars"].copy()
    chars[90_000:90_700] |= 0x20
    case = dict(case, chars=chars)
    scanner.set_engine(capi.ENGINE_AUTO)
    scanner.set_motifs(case["P"], case["col_len"], case["thr"])
    codes, zm, has_zero = capi.pack_ascii(chars, lower)
    assert has_zero == (lower == capi.LOWER_ZERO)
    n = len(chars)
    scanner.submit_packed(0, codes, zm if has_zero else None, n, n, case["frag_start"][1:])
    scanner.submit_ascii(1, chars, frag_starts=case["frag_start"][1:], lower=lower)
    hp, _ = scanner.collect(0)
    ha, _ = scanner.collect(1)
    want = _oracle_hits(case, lower_fold=(lower == capi.LOWER_FOLD))
    _assert_same(hp, *want)
    _assert_same(ha, *want)


@pytest.mark.parametrize("engine", [capi.ENGINE_GATHER, capi.ENGINE_TENSOR])
def test_hit_buffer_overflow_regrows(engine):
    """Thresholds so low that almost every window is an occurrence: far more hits than max_hits."""
    s = capi.Scanner(0, max_block_nt=1 << 20, max_hits=1024)
    try:
        case = util.random_case(51, n_motifs=4, n_nt=40_000, len_range=(5, 9))
        thr = np.full(len(case["thr"]), -1000.0, dtype=np.float32)
        s.set_engine(engine)
        s.set_motifs(case["P"], case["col_len"], thr)
        hits, _ = s.scan(case["chars"], case["frag_start"][1:])
        _assert_same(hits, *_oracle_hits(dict(case, thr=thr)))
        assert len(hits) > 100_000
    finally:
        s.close()


@pytest.mark.parametrize("engine", [capi.ENGINE_GATHER, capi.ENGINE_TENSOR])
def test_compact_hit_records(engine):
    """b200scan_set_hit_format(B200SCAN_HITS_12) / b200scan_collect12: the same occurrences as 12-byte records -- also
    through the regrow path (max_hits = 1024) -- and the state errors of mixing the two formats."""
    s = capi.Scanner(0, max_block_nt=1 << 20, max_hits=1024)
    try:
        case = util.random_case(71, n_motifs=12, n_nt=150_000, lower=True)
        s.set_engine(engine)
        s.set_motifs(case["P"], case["col_len"], case["thr"])
        want = _oracle_hits(case)
        s.set_hit_format(capi.HITS_12)
        h12, _ = s.scan(case["chars"], case["frag_start"][1:])
        assert h12.dtype == capi.HIT12_DTYPE and h12.dtype.itemsize == 12 and len(h12) > 1024
        _assert_same(h12, *want)
        # both slots in flight in the compact format, then back to 16-byte records
        s.submit_ascii(0, case["chars"], frag_starts=case["frag_start"][1:])
        fs = case["frag_start"][1:]
        s.submit_ascii(1, case["chars"][:70_029], n_payload=70_000, frag_starts=fs[fs < 70_029])     # payload + halo of maxLen - 1
        with pytest.raises(capi.ScanError):
            s.set_hit_format(capi.HITS_16)                   # a block is in flight
        with pytest.raises(capi.ScanError) as ei:
            s.collect(0, fmt=capi.HITS_16)                   # submitted under the 12-byte format
        assert ei.value.code != 0
        a, _ = s.collect(0)
        b, _ = s.collect(1)
        _assert_same(a, *want)
        k = want[0] < 70_000
        _assert_same(b, want[0][k], want[1][k], want[2][k])
        s.set_hit_format(capi.HITS_16)
        h16, _ = s.scan(case["chars"], case["frag_start"][1:])
        assert h16.dtype == capi.HIT_DTYPE
        _assert_same(h16, *want)
        with pytest.raises(capi.ScanError):
            s.set_hit_format(13)
    finally:
        s.close()


@pytest.mark.parametrize("engine", [capi.ENGINE_GATHER, capi.ENGINE_TENSOR])
def test_ordered_hit_records(engine):
    """B200SCAN_HITS_8 / b200scan_collect8 (csrc/order.cuh): the device returns the block's occurrences already in
    (position, column) order as 8-byte records + a bucket index -- compared RECORD BY RECORD, without sorting, with the
    oracle's list (which is in that order); through the regrow path (max_hits = 1024), with all three slots in flight, for
    an empty block, for a block whose last bucket is partial, and the state errors of mixing the formats."""
    s = capi.Scanner(0, max_block_nt=1 << 20, max_hits=1024)
    try:
        case = util.random_case(73, n_motifs=12, n_nt=150_001, lower=True)
        s.set_engine(engine)
        s.set_motifs(case["P"], case["col_len"], case["thr"])
        want = _oracle_hits(case)
        s.set_hit_format(capi.HITS_8)
        s.submit_ascii(0, case["chars"], frag_starts=case["frag_start"][1:])
        h8, bstart, t = s.collect8(0)
        assert h8.dtype.itemsize == 8 and len(h8) == len(want[0]) > 1024 and t["order_ms"] > 0
        assert len(bstart) == (150_001 + 255) // 256 + 1 and bstart[0] == 0 and bstart[-1] == len(h8)
        assert np.all(np.diff(bstart.astype(np.int64)) >= 0)
        h = capi.expand_hits8(h8, bstart)
        assert np.array_equal(h["pos"], want[0]) and np.array_equal(h["col"], want[1])          # in order: no sort here
        assert np.array_equal(h["score"].view(np.uint32), want[2].view(np.uint32))
        # three blocks in flight, of different sizes (the ordering scratch buffers are shared by the slots)
        fs = case["frag_start"][1:]
        s.submit_ascii(0, case["chars"], frag_starts=fs)
        s.submit_ascii(1, case["chars"][:70_029], n_payload=70_000, frag_starts=fs[fs < 70_029])
        s.submit_ascii(2, case["chars"][:300], n_payload=256, frag_starts=fs[fs < 300])
        with pytest.raises(capi.ScanError):
            s.set_hit_format(capi.HITS_12)                   # blocks in flight
        with pytest.raises(capi.ScanError):
            s.collect(0, fmt=capi.HITS_12)                   # submitted under the 8-byte format
        for slot, lim in ((2, 256), (0, 1 << 30), (1, 70_000)):
            got, _ = s.collect(slot)                         # Scanner.collect expands ordered records itself
            k = want[0] < lim
            assert np.array_equal(got["pos"], want[0][k]) and np.array_equal(got["col"], want[1][k])
            assert np.array_equal(got["score"].view(np.uint32), want[2][k].view(np.uint32))
        # empty block, then a dense one: every window of 4 columns is an occurrence (many hits per position and bucket)
        s.submit_ascii(0, np.zeros(0, np.uint8))
        h8, bstart, _ = s.collect8(0)
        assert len(h8) == 0 and len(bstart) == 1 and bstart[0] == 0
        dense = util.random_case(51, n_motifs=4, n_nt=40_000, len_range=(5, 9))
        thr = np.full(len(dense["thr"]), -1000.0, dtype=np.float32)
        s.set_motifs(dense["P"], dense["col_len"], thr)
        got, _ = s.scan(dense["chars"], dense["frag_start"][1:])
        wd = _oracle_hits(dict(dense, thr=thr))
        assert len(got) > 100_000 and np.array_equal(got["pos"], wd[0]) and np.array_equal(got["col"], wd[1])
        assert np.array_equal(got["score"].view(np.uint32), wd[2].view(np.uint32))
        s.set_hit_format(capi.HITS_16)
        h16, _ = s.scan(dense["chars"], dense["frag_start"][1:])
        _assert_same(h16, *wd)
    finally:
        s.close()


def test_double_buffered_slots_and_state_errors(scanner):
    a, b = util.random_case(61, n_motifs=8, n_nt=80_000), util.random_case(62, n_motifs=8, n_nt=70_000)
    scanner.set_engine(capi.ENGINE_AUTO)
    scanner.set_motifs(a["P"], a["col_len"], a["thr"])
    scanner.submit_ascii(0, a["chars"], frag_starts=a["frag_start"][1:])
    scanner.submit_ascii(1, b["chars"], frag_starts=b["frag_start"][1:])
    with pytest.raises(capi.ScanError):
        scanner.submit_ascii(0, a["chars"])                  # slot busy
    h1, _ = scanner.collect(1)
    h0, _ = scanner.collect(0)
    _assert_same(h0, *_oracle_hits(a))
    _assert_same(h1, *_oracle_hits(dict(b, P=a["P"], col_len=a["col_len"], thr=a["thr"])))
    with pytest.raises(capi.ScanError):
        scanner.collect(0)                                   # nothing in flight
    with pytest.raises(capi.ScanError):
        scanner.submit_ascii(0, a["chars"], frag_starts=np.array([5, 5], dtype=np.uint64))   # not ascending


def test_pipelined_lower_case_block_keeps_zero_mask(scanner):
    """A lower-case block submitted on slot 1 while slot 0's kernels are still running: its upload, counter reset and
    packing run on the upload stream, and the has_zero flag the pack kernel raises must survive until the scan of that
    block (its tensor instance zeroes the rows of masked characters; without the flag they would score as upper case)."""
    big = util.random_case(61, n_motifs=60, n_nt=24_000_000, len_range=(6, 14), with_gaps=False)
    small = util.random_case(62, n_motifs=60, n_nt=300_000, len_range=(6, 14), lower=True)
    for c in (big, small):
        c["P"], c["col_len"] = big["P"], big["col_len"]
        c["thr"] = np.maximum(big["thr"], 8.0).astype(np.float32)
    scanner.set_engine(capi.ENGINE_AUTO)
    scanner.set_motifs(big["P"], big["col_len"], big["thr"])
    for _ in range(3):
        scanner.submit_ascii(0, big["chars"])
        scanner.submit_ascii(1, small["chars"], frag_starts=small["frag_start"][1:])
        h0, t0 = scanner.collect(0)
        h1, t1 = scanner.collect(1)
        assert t0["engine_used"] == capi.ENGINE_TENSOR and t1["engine_used"] == capi.ENGINE_TENSOR
        _assert_same(h1, *_oracle_hits(small))
    assert len(h0) > 1000


def test_engines_agree_at_scale_and_properties(scanner):
    """32 Mnt x 400 columns (1.3e10 scores): too big for the oracle's full scan, so check size-independent
    properties: both engines return the identical hit list; every reported score re-verifies against the
    oracle at that (position, column); every hit clears its threshold and lies inside one fragment; shifting
    the block start by k moves every hit by k (translation invariance)."""
    case = util.random_case(71, n_motifs=200, n_nt=1 << 25, len_range=(5, 30))
    for c in range(len(case["thr"])):
        case["thr"][c] = max(case["thr"][c], 11.0)
    res = {}
    for engine, acc in ((capi.ENGINE_GATHER, 0), (capi.ENGINE_TENSOR, 16), (capi.ENGINE_TENSOR, 32)):
        scanner.set_engine(engine)
        scanner.set_tensor_accumulator(acc)
        scanner.set_motifs(case["P"], case["col_len"], case["thr"])
        res[acc], t = scanner.scan(case["chars"], case["frag_start"][1:])
        res[acc] = _sorted(res[acc])
    scanner.set_tensor_accumulator(0)
    g, tcs = res[0], res[16]
    assert len(g) > 1000 and np.array_equal(g, tcs) and np.array_equal(g, res[32])
    want = O.score_at(bytes(case["chars"]), case["P"], case["col_len"], g["pos"], g["col"])
    assert np.array_equal(want.view(np.uint32), g["score"].view(np.uint32))
    assert np.all(g["score"] >= case["thr"][g["col"]])
    fi = np.searchsorted(case["frag_start"], g["pos"], side="right")
    end = np.append(case["frag_start"], len(case["chars"]))[fi]
    assert np.all(g["pos"] + case["col_len"][g["col"]].astype(np.uint64) <= end)
    k = 4099
    scanner.set_engine(capi.ENGINE_TENSOR)
    shifted, _ = scanner.scan(case["chars"][k:], (case["frag_start"][case["frag_start"] > k] - np.uint64(k)))
    shifted = _sorted(shifted)
    # hits whose window lies in a fragment that started before k may appear (their fragment is cut open); others must map 1:1
    first_frag_after = case["frag_start"][case["frag_start"] > k].min()
    keep = g[g["pos"] >= first_frag_after].copy(); keep["pos"] -= np.uint64(k)
    assert np.array_equal(keep, shifted[shifted["pos"] >= first_frag_after - k])


def test_full_size_bench_workload_properties(tmp_path):
    """BASELINE.json configs[1] at full size (bench.py's workload: 1800 columns x 100 Mbp = 1.8e11 scores), through
    size-independent properties: the tensor path and the exact gather-add engine return the identical hit list; a random
    sample of hits re-verifies bit-exactly against the oracle; every hit clears its threshold; a window-aligned slice of
    the block scanned on its own returns exactly the hits of that slice (chunking invariance, the multi-GPU shard rule)."""
    import bench
    ms, P, col_len, thr, seq, bg = bench.build_inputs(str(tmp_path), 100_000_000, 0)
    assert len(col_len) == 1800
    sc = capi.Scanner(0, max_block_nt=len(seq) + 64, max_hits=1 << 25)
    res = {}
    for engine in (capi.ENGINE_TENSOR, capi.ENGINE_GATHER):
        sc.set_engine(engine)
        sc.set_motifs(P, col_len, thr)
        h, t = sc.scan(seq)
        assert t["engine_used"] == engine
        res[engine] = _sorted(h)
    tc, g = res[capi.ENGINE_TENSOR], res[capi.ENGINE_GATHER]
    assert len(g) > 10_000_000 and np.array_equal(tc, g)
    assert np.all(g["score"] >= thr[g["col"]])
    assert np.all(g["pos"] + col_len[g["col"]].astype(np.uint64) <= len(seq))
    pick = np.sort(np.random.default_rng(5).choice(len(g), 200_000, replace=False))
    want = O.score_at(bytes(seq), P, col_len, g["pos"][pick], g["col"][pick])
    assert np.array_equal(want.view(np.uint32), g["score"][pick].view(np.uint32))
    lo, n_pay, halo = 37_000_001, 9_000_000, int(col_len.max()) - 1
    sc.set_engine(capi.ENGINE_TENSOR)
    sc.set_motifs(P, col_len, thr)
    part, _ = sc.scan(seq[lo:lo + n_pay + halo], None, n_pay)
    part = _sorted(part)
    ref = g[(g["pos"] >= lo) & (g["pos"] < lo + n_pay)].copy()
    ref["pos"] -= np.uint64(lo)
    assert np.array_equal(part, ref)
    sc.close()


def test_sharded_scan_matches_single_pass(scanner):
    """The multi-GPU decomposition (chunks + halo, round-robin over ranks, host merge) on one device."""
    case = util.random_case(81, n_motifs=20, n_nt=500_000)
    halo = int(case["col_len"].max()) - 1
    scanner.set_engine(capi.ENGINE_AUTO)
    scanner.set_motifs(case["P"], case["col_len"], case["thr"])
    shards = shard.plan_shards(len(case["chars"]), world=4, halo=halo, chunk=70_001)
    parts = []
    for s in shards:
        block = case["chars"][s.start:s.start + s.n_total]
        h, _ = scanner.scan(block, shard.local_frag_starts(case["frag_start"], s), s.n_payload)
        parts.append(h)
    merged = shard.merge_hits(parts, shards)
    pos, col, sc = _oracle_hits(case)
    assert np.array_equal(merged["pos"], pos) and np.array_equal(merged["col"], col)
    assert np.array_equal(merged["score"].view(np.uint32), sc.view(np.uint32))


def test_cli_end_to_end_example(golden, tmp_path):
    """blamm-b200 dict / hist / scan on the reference's example: occurrences.txt == the reference's, sorted."""
    cli = os.path.join(lib_dir(), "blamm-b200")
    work = tmp_path / "ex"
    shutil.copytree(os.path.join(golden, "example"), work)
    for f in os.listdir(work):
        if os.path.isfile(work / f) and (f.startswith(("hist_", "occ_", "refdump_", "PWM")) or f.endswith(".dict")):
            os.remove(work / f)
    env = dict(os.environ, BLAMM_B200_CHUNK="30000")
    for args in (["dict", "sequences.mf"], ["hist", "motifs.jaspar", "sequences.mf"]):
        subprocess.run([cli] + args, cwd=work, check=True, stdout=subprocess.DEVNULL)
    for mode_key, flags in (("pt_rc", ["-rc", "-pt", "0.0001"]), ("pt_fwd", ["-pt", "0.0001"]), ("rt_rc", ["-rc"]),
                            ("at_rc", ["-rc", "-at", "9.5"])):
        r = subprocess.run([cli, "scan"] + flags + ["motifs.jaspar", "sequences.mf"], cwd=work, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        got = sorted(open(work / "occurrences.txt").read().splitlines(True))
        assert got == open(os.path.join(golden, "example", "occ_%s.txt" % mode_key)).read().splitlines(True)
        assert ("Wrote %d matches" % len(got)) in r.stdout
        if mode_key == "pt_rc":
            assert open(work / "PWMthresholds.txt").read() == open(os.path.join(golden, "example", "PWMthresholds_pt_rc.txt")).read()


def test_config4_many_columns_absolute_threshold(tmp_path):
    """BASELINE.json configs[3] in shape: 10,000 PWMs of length 6-30 with -rc (20,000 columns, 80 column tiles) and an
    absolute threshold, at 2 Mnt.  The oracle scans a prefix in full; beyond it the tensor path must equal the exact
    gather-add engine hit for hit, and a sample of hits re-verifies bit-exactly against the oracle."""
    mf = str(tmp_path / "m10k.jaspar")
    synth.make_jaspar_like(mf, 10000, 77, uniform_len=(6, 30))
    seq = synth.random_acgt(2_000_000, 99)
    ms = capi.MotifSet(mf, revcompl=True)
    P, col_len, is_rc = ms.generate_matrix(synth.counts_of(seq))
    thr = ms.thresholds("at", 12.0)
    assert len(col_len) == 20000 and int(col_len.min()) == 6 and int(col_len.max()) == 30
    sc = capi.Scanner(0, max_block_nt=len(seq) + 64, max_hits=1 << 23)
    try:
        res = {}
        for engine in (capi.ENGINE_TENSOR, capi.ENGINE_GATHER):
            sc.set_engine(engine)
            sc.set_motifs(P, col_len, thr)
            h, t = sc.scan(seq)
            assert t["engine_used"] == engine
            res[engine] = _sorted(h)
        tc, g = res[capi.ENGINE_TENSOR], res[capi.ENGINE_GATHER]
        assert len(g) > 100_000 and np.array_equal(tc, g)
        assert np.all(g["score"] >= thr[g["col"]])
        assert np.all(g["pos"] + col_len[g["col"]].astype(np.uint64) <= len(seq))
        pick = np.sort(np.random.default_rng(6).choice(len(g), 100_000, replace=False))
        want = O.score_at(bytes(seq), P, col_len, g["pos"][pick], g["col"][pick])
        assert np.array_equal(want.view(np.uint32), g["score"][pick].view(np.uint32))
        # full oracle scan of a prefix (every column): identical set, bit-identical scores
        n = 6000
        pos, col, score = O.scan_stream(bytes(seq[:n]), np.zeros(1, np.uint64), P, col_len, thr)
        sc.set_engine(capi.ENGINE_TENSOR)
        h, _ = sc.scan(seq[:n])
        _assert_same(h, pos, col, score)
        assert np.array_equal(_sorted(h), tc[tc["pos"] + col_len[tc["col"]].astype(np.uint64) <= n])
    finally:
        sc.close()


@pytest.mark.parametrize("handover", ["packed", "ascii"])
def test_config3_many_groups_cli(tmp_path, handover):
    """(`handover`: the CLI's default -- chunks packed to 2 bits by the parser threads, b200scan_submit_packed -- and
    BLAMM_B200_ASCII=1 -- characters sent, packed on the device; the packed run is repeated with -s, lower case folded.)
    BASELINE.json configs[2] in shape: several manifest groups with distinct backgrounds (the matrix P and the p-value
    thresholds are rebuilt per group), more than one FASTA file per group, N runs and soft-masked stretches, scanned by the
    CLI in small chunks with a parallel reader -- against the oracle's restatement of the whole `blamm scan`."""
    cli = os.path.join(lib_dir(), "blamm-b200")
    work = tmp_path / "c3"
    os.makedirs(work / "hist")
    synth.make_jaspar_like(str(work / "motifs.jaspar"), 40, 31)
    rng = np.random.default_rng(8)
    manifest = []
    for gidx, gc in enumerate((0.36, 0.41, 0.44, 0.48, 0.52)):
        probs = ((1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2)
        for f in range(1 + gidx % 2):
            seq = synth.random_acgt(150_000 + 10_000 * gidx, 100 + 10 * gidx + f, probs)
            for _ in range(4):
                a = int(rng.integers(0, len(seq) - 5000))
                seq[a:a + int(rng.integers(1, 3000))] = ord("N")
                b = int(rng.integers(0, len(seq) - 5000))
                seq[b:b + int(rng.integers(1, 4000))] |= 0x20
            cut = len(seq) // 3
            name = "g%d_%d.fa" % (gidx, f)
            synth.write_fasta(str(work / name), [("g%dchr%d" % (gidx, 2 * f + 1), seq[:cut]), ("g%dchr%d extra" % (gidx, 2 * f + 2), seq[cut:])])
            manifest.append("group%d\t%s\n" % (gidx, name))
    open(work / "seq.mf", "w").write("".join(manifest))
    env = dict(os.environ, BLAMM_B200_CHUNK="50000", BLAMM_B200_INGEST_THREADS="4", BLAMM_B200_ASCII="1" if handover == "ascii" else "0")
    for args in (["dict", "seq.mf"], ["hist", "-H", "hist", "motifs.jaspar", "seq.mf"]):
        subprocess.run([cli] + args, cwd=work, env=env, check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run([cli, "scan", "-rc", "-pt", "0.0005", "-H", "hist", "motifs.jaspar", "seq.mf"], cwd=work, env=env,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want, details = O.scan("motifs.jaspar", "seq.mf", "pt", 0.0005, True, histdir="hist", base_dir=str(work))
    assert len(details) == 5 and all(len(d["pos"]) > 500 for d in details)
    # distinct backgrounds -> distinct thresholds per group
    assert not np.array_equal(details[0]["thr"], details[4]["thr"])
    got = sorted(open(work / "occurrences.txt").read().splitlines(True))
    assert got == sorted(want)
    if handover == "packed":
        r = subprocess.run([cli, "scan", "-s", "-rc", "-pt", "0.0005", "-H", "hist", "-o", "occ_s.txt", "motifs.jaspar", "seq.mf"], cwd=work,
                           env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        want_s, _ = O.scan("motifs.jaspar", "seq.mf", "pt", 0.0005, True, histdir="hist", base_dir=str(work), lower_fold=True)
        got_s = sorted(open(work / "occ_s.txt").read().splitlines(True))
        assert got_s == sorted(want_s) and got_s != got


@pytest.mark.parametrize("lower", [capi.LOWER_ZERO, capi.LOWER_FOLD])
def test_empirical_histogram_matches_oracle(scanner, lower):
    """b200scan_hist_* (the `blamm hist -e` epilogue on the GPU) against the oracle: every bin identical, including
    fragment boundaries, lower-case handling, several blocks with a halo and more columns than one shared-memory tile."""
    case = util.random_case(91, n_motifs=60, n_nt=150_000, len_range=(5, 33), lower=True)
    P, col_len = case["P"], case["col_len"]
    mm = [O.max_min_score(np.ascontiguousarray(P[c, :4 * int(col_len[c])])) for c in range(len(col_len))]
    mx = np.array([a for a, b in mm], np.float32); mn = np.array([b for a, b in mm], np.float32)
    bins = 250
    want = O.empirical_hist(bytes(case["chars"]), case["frag_start"], P, col_len, mn, mx, bins, lower_fold=(lower == capi.LOWER_FOLD))
    scanner.set_engine(capi.ENGINE_AUTO)
    scanner.set_motifs(P, col_len, np.zeros(len(col_len), np.float32))
    scanner.hist_begin(mn, mx, bins)
    halo = int(col_len.max()) - 1
    for s in shard.plan_shards(len(case["chars"]), world=1, halo=halo, chunk=40_003):
        scanner.hist_block(case["chars"][s.start:s.start + s.n_total], shard.local_frag_starts(case["frag_start"], s), s.n_payload, lower=lower)
    got = scanner.hist_read()
    assert got.shape == want.shape and int(got.sum()) == int(want.sum()) > 0
    assert np.array_equal(got, want)


def test_cli_empirical_histograms_example(golden, tmp_path):
    """`blamm-b200 hist -e` writes the same .dat files as the reference's `blamm hist -e` on the example."""
    cli = os.path.join(lib_dir(), "blamm-b200")
    work = tmp_path / "ex"
    shutil.copytree(os.path.join(golden, "example"), work)
    out = work / "he"
    out.mkdir()
    r = subprocess.run([cli, "hist", "-e", "-H", "he", "motifs.jaspar", "sequences.mf"], cwd=work, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref = os.path.join(golden, "example", "hist_e")
    files = sorted(f for f in os.listdir(ref) if f.endswith(".dat"))
    assert len(files) == 4
    for f in files:
        assert (out / f).read_bytes() == open(os.path.join(ref, f), "rb").read(), f


def test_block_offsets_beyond_2_to_31():
    """One block of 2.3e9 characters (the ABI takes up to 2^32 - 2^28): window positions, code-word offsets and bucket indices
    pass 2^31.  The block is a 1,000,003-character segment repeated 2300 times, so the exact answer is periodic: the oracle scans
    two periods once, and every period of the block -- in particular those that lie beyond position 2^31 -- must return exactly
    its hits (position + k * period), in order, through the host packer, b200scan_submit_packed and the ordered 8-byte records."""
    period, reps = 1_000_003, 2300
    case = util.random_case(91, n_motifs=6, n_nt=period, len_range=(8, 20), with_gaps=False)
    thr = (case["thr"] + np.float32(3.0)).astype(np.float32)                 # a few hundred hits per period
    seg = case["chars"]
    two = np.concatenate([seg, seg])
    pos, col, sc = O.scan_stream(bytes(two), np.zeros(1, np.uint64), case["P"], case["col_len"], thr)
    k = pos < period
    pos, col, sc = pos[k], col[k], sc[k]
    assert 50 < len(pos) < 20000
    block = np.tile(seg, reps)
    n = len(block)
    assert n > (1 << 31) + (1 << 26)
    codes, zm, has_zero = capi.pack_ascii(block)
    del block
    assert not has_zero
    s = capi.Scanner(0, max_block_nt=n + 64, max_hits=1 << 22)
    try:
        s.set_engine(capi.ENGINE_AUTO)
        s.set_hit_format(capi.HITS_8)
        s.set_motifs(case["P"], case["col_len"], thr)
        s.submit_packed(0, codes, None, n, n)
        h8, bstart, t = s.collect8(0)
        assert t["engine_used"] == capi.ENGINE_TENSOR
        h = capi.expand_hits8(h8, bstart)
        # the last period has no successor to wrap into: count what the oracle finds in a lone period for it
        p1, c1, s1 = O.scan_stream(bytes(seg), np.zeros(1, np.uint64), case["P"], case["col_len"], thr)
        assert len(h) == (reps - 1) * len(pos) + len(p1)
        hp = h["pos"].astype(np.int64)
        assert np.all(np.diff(hp) >= 0)
        for rep in (0, 1073, 2147, 2148, reps - 2):                          # 2147 / 2148: the periods around position 2^31
            m = (hp >= rep * period) & (hp < (rep + 1) * period)
            assert np.array_equal(hp[m] - rep * period, pos.astype(np.int64)) and np.array_equal(h["col"][m], col)
            assert np.array_equal(h["score"][m].view(np.uint32), sc.view(np.uint32))
        m = hp >= (reps - 1) * period
        assert np.array_equal(hp[m] - (reps - 1) * period, p1.astype(np.int64)) and np.array_equal(h["col"][m], c1)
    finally:
        s.close()


def test_cli_splits_chunks_too_dense_for_the_device_buffers(tmp_path):
    """Thresholds so low that a third of all windows are occurrences, and a device budget of 30,000 hit records
    (B200SCAN_HIT_BUDGET): b200scan_collect8 refuses the chunks with B200SCAN_ENOMEM and the CLI scores them in halves,
    recursively (cli.cpp: scanSplit), still writing every chunk in stream order.  The occurrence file must equal the oracle's
    whole-scan restatement, line for line after sorting -- and, per sequence, already be in position order."""
    cli = os.path.join(lib_dir(), "blamm-b200")
    work = str(tmp_path)
    rng = np.random.default_rng(3)
    synth.make_jaspar_like(os.path.join(work, "motifs.jaspar"), 6, 11, uniform_len=(6, 10))
    seq = synth.random_acgt(300_000, 12)
    seq[100_000:100_050] = ord("N")
    synth.write_fasta(os.path.join(work, "a.fa"), [("s1", seq[:180_000]), ("s2", seq[180_000:])])
    open(os.path.join(work, "sequences.mf"), "w").write("g\ta.fa\n")
    subprocess.run([cli, "dict", "sequences.mf"], cwd=work, check=True, stdout=subprocess.DEVNULL)
    env = dict(os.environ, B200SCAN_HIT_BUDGET="30000", BLAMM_B200_CHUNK="70000")
    r = subprocess.run([cli, "scan", "-rc", "-at", "-6", "-t", "4", "motifs.jaspar", "sequences.mf"], cwd=work, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = open(os.path.join(work, "occurrences.txt")).read().splitlines(True)
    want, _ = O.scan("motifs.jaspar", "sequences.mf", mode="at", value=-6.0, revcompl=True, base_dir=work)
    assert len(got) == len(want) > 300_000
    assert sorted(got) == sorted(want)
    starts = {}
    for l in got:
        c = l.split("\t")
        assert int(c[3]) >= starts.get(c[0], 0)
        starts[c[0]] = int(c[3])
    # without the budget the same run takes the normal path and writes the same file
    r = subprocess.run([cli, "scan", "-rc", "-at", "-6", "-t", "4", "-o", "plain.txt", "motifs.jaspar", "sequences.mf"], cwd=work, capture_output=True, text=True)
    assert r.returncode == 0 and open(os.path.join(work, "plain.txt")).read().splitlines(True) == got


@pytest.mark.parametrize("forced", [False, True])
def test_rescorer_tiles_beyond_the_shared_memory_budget(forced, monkeypatch):
    """The fused rescorer keeps a column tile's FP32 weights in shared memory; a tile that does not fit (256 columns of more than
    ~38 positions) is scored from global memory by the same code.  Natural case: 260 columns of length 48..64; forced case: a budget
    of 64 positions (B200SCAN_FUSE_MAXW) puts every tile of an ordinary set on that path.  Bit-exact against the oracle, with a
    soft-masked stretch (the masked instance takes the same path)."""
    if forced:
        monkeypatch.setenv("B200SCAN_FUSE_MAXW", "64")
        case = util.random_case(81, n_motifs=60, n_nt=400_000, len_range=(5, 30), lower=True)
    else:
        case = util.random_case(82, n_motifs=130, n_nt=200_000, len_range=(48, 64), lower=True)
        assert len(case["col_len"]) == 260 and int(np.sort(case["col_len"])[-256:].sum()) > 9728
        # (such long random motifs never reach the usual thresholds: a low absolute one gives 25,214 occurrences with lower case
        #  scored like upper case, and millions where the masked stretch contributes zero -- the regrow path on top)
        case["thr"] = np.full(len(case["thr"]), -20.0, dtype=np.float32)
    s = capi.Scanner(0, max_block_nt=1 << 20, max_hits=1 << 20)
    try:
        s.set_engine(capi.ENGINE_TENSOR)
        s.set_motifs(case["P"], case["col_len"], case["thr"])
        for lower in (capi.LOWER_ZERO, capi.LOWER_FOLD):
            hits, t = s.scan(case["chars"], case["frag_start"][1:], lower=lower)
            assert t["engine_used"] == capi.ENGINE_TENSOR
            want = _oracle_hits(case, lower_fold=(lower == capi.LOWER_FOLD))
            assert len(want[0]) > 0
            _assert_same(hits, *want)
    finally:
        s.close()


def test_cli_reports_a_stale_dictionary_instead_of_aborting(tmp_path):
    """A FASTA file that changed after `dict` (one record more than the dictionary knows): the formatting threads meet a record
    index without a name.  The run must end with an error message and exit code 1 (the reference's behaviour for every failure,
    blstools.cpp:93-101), not with std::terminate from a pool thread."""
    cli = os.path.join(lib_dir(), "blamm-b200")
    work = str(tmp_path)
    synth.make_jaspar_like(os.path.join(work, "motifs.jaspar"), 8, 5, uniform_len=(6, 9))
    seq = synth.random_acgt(90_000, 21)
    synth.write_fasta(os.path.join(work, "a.fa"), [("s1", seq[:45_000]), ("s2", seq[45_000:])])
    open(os.path.join(work, "sequences.mf"), "w").write("g\ta.fa\n")
    subprocess.run([cli, "dict", "sequences.mf"], cwd=work, check=True, stdout=subprocess.DEVNULL)
    synth.write_fasta(os.path.join(work, "a.fa"), [("s1", seq[:30_000]), ("s2", seq[30_000:60_000]), ("s3", seq[60_000:])])
    r = subprocess.run([cli, "scan", "-rc", "-at", "2", "motifs.jaspar", "sequences.mf"], cwd=work, capture_output=True, text=True)
    assert r.returncode == 1, (r.returncode, r.stderr[-500:])
    assert r.stderr.strip() != "" and "bye" not in r.stdout


@pytest.mark.parametrize("kernel", [3, 2, 1, 0])
def test_histogram_kernels_agree_with_oracle(monkeypatch, kernel):
    """Every histogram kernel of the library (B200SCAN_HIST_KERNEL, read by b200scan_create: 3 = position-row weights with the lanes
    of a warp counting different columns, the default; 2 = the same with match-aggregated counts; 1 / 0 = the round-1 kernel with /
    without aggregated counts) against the oracle, bin for bin: 150 columns of 5 to 64 positions (several shared-memory tiles, the
    longest column the build supports), lower case under both rules, fragments, several blocks with a halo."""
    monkeypatch.setenv("B200SCAN_HIST_KERNEL", str(kernel))
    case = util.random_case(92, n_motifs=150, n_nt=120_000, len_range=(5, 64), lower=True)
    P, col_len = case["P"], case["col_len"]
    mm = [O.max_min_score(np.ascontiguousarray(P[c, :4 * int(col_len[c])])) for c in range(len(col_len))]
    mx = np.array([a for a, b in mm], np.float32); mn = np.array([b for a, b in mm], np.float32)
    bins = 250
    halo = int(col_len.max()) - 1
    s = capi.Scanner(0, max_block_nt=1 << 20, max_hits=1 << 16)
    try:
        s.set_motifs(P, col_len, np.zeros(len(col_len), np.float32))
        for lower in (capi.LOWER_ZERO, capi.LOWER_FOLD):
            want = O.empirical_hist(bytes(case["chars"]), case["frag_start"], P, col_len, mn, mx, bins, lower_fold=(lower == capi.LOWER_FOLD))
            s.hist_begin(mn, mx, bins)
            for sh in shard.plan_shards(len(case["chars"]), world=1, halo=halo, chunk=50_001):
                s.hist_block(case["chars"][sh.start:sh.start + sh.n_total], shard.local_frag_starts(case["frag_start"], sh), sh.n_payload, lower=lower)
            got = s.hist_read()
            assert got.shape == want.shape and int(got.sum()) == int(want.sum()) > 0
            assert np.array_equal(got, want)
    finally:
        s.close()
