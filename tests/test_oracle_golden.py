"""The oracle (oracle/oracle.c + oracle.py) against fixtures produced by the compiled reference
(tests/golden/make_golden.py): occurrence text, and -- through refdump -- P, thresholds and hit scores
bit for bit.  This is what pins the oracle (the reference ships no golden vectors of its own)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle.refdump_io import read_refdump
from tests import util


def _check_case(case_dir, mode_key, bit_exact=True):
    lines, det = util.oracle_case(case_dir, mode_key)
    want = open(os.path.join(case_dir, "occ_%s.txt" % mode_key)).read().splitlines(True)
    if bit_exact:
        assert sorted(lines) == want
    else:   # the printed 6-digit score may flip in its last digit when the reference's own sum moved by an ulp
        def split(ls):
            rows = sorted(l.split("\t") for l in ls)
            return [r[:5] + r[6:] for r in rows], np.array([float(r[5]) for r in rows])
        (mk, ms), (wk, ws) = split(lines), split(want)
        assert sorted(mk) == sorted(wk)
    dump = read_refdump(os.path.join(case_dir, "refdump_%s.bin" % mode_key))
    assert len(dump) == len(det)
    for d, r in zip(det, dump):
        assert r["name"] == d["species"].name
        names = [c["name"] for c in r["cols"]]
        # same multiset of (name, strand, length); column order may differ between equal-length motifs
        assert sorted((m.name, m.revcomp, len(m)) for m in d["motifs"]) == sorted((c["name"], c["rc"], c["len"]) for c in r["cols"])
        # map oracle columns -> reference columns by (name, strand)
        ref_col = {(c["name"], c["rc"]): j for j, c in enumerate(r["cols"])}
        cmap = np.array([ref_col[(m.name, m.revcomp)] for m in d["motifs"]])
        for c, m in enumerate(d["motifs"]):
            assert np.array_equal(d["P"][c].view(np.uint32), r["P"][cmap[c]].view(np.uint32)), "P differs for " + m.name
            assert np.float32(d["thr"][c]).view(np.uint32) == np.float32(r["cols"][cmap[c]]["thr"]).view(np.uint32)
        util.compare_with_refdump(d, r, blas_bit_exact=bit_exact)


@pytest.mark.parametrize("mode_key", ["pt_rc", "pt_fwd", "rt_rc", "at_rc"])
def test_example(golden, mode_key):
    _check_case(os.path.join(golden, "example"), mode_key)


def test_example_md5_of_survey(golden):
    # sorted occurrences of `scan -rc -pt 0.0001` on the example
    import hashlib
    data = open(os.path.join(golden, "example", "occ_pt_rc.txt"), "rb").read()
    assert hashlib.md5(data).hexdigest() == "d718476656e56c92414235792b28d534"
    assert data.count(b"\n") == 124


@pytest.mark.parametrize("mode_key", ["pt_rc", "pt_fwd", "rt_rc", "at_rc", "at_low"])
def test_edge_cases(golden, mode_key):
    _check_case(os.path.join(golden, "edge"), mode_key, bit_exact=False)


def test_synth2m_bits(golden, tmp_path):
    d = util.materialise_synth2m(str(tmp_path))
    lines, det = O.scan("motifs.jaspar", "sequences.mf", "pt", 1e-4, True, histdir=".", base_dir=d)
    r = read_refdump(os.path.join(d, "refdump_pt_rc.bin"))[0]
    dd = det[0]
    assert len(dd["pos"]) == len(r["hits"]) == 21181
    util.compare_with_refdump(dd, r, blas_bit_exact=False)      # motifs up to 35 long: BLAS re-associates


def test_hist_files_match_reference(golden):
    for case in ("example", "edge"):
        d = os.path.join(golden, case)
        species = O.load_dict(os.path.join(d, "sequences.mf.dict"))
        motifs = O.load_jaspar(os.path.join(d, "motifs.jaspar"))
        for sp in species:
            for m in motifs:
                pwm = O.pwm_of(m, sp.counts)
                txt = O.hist_dat_text(pwm, O.theoretical_hist(pwm, sp.counts))
                assert txt == open(os.path.join(d, "hist_%s_%s.dat" % (sp.name, m.name))).read(), (case, sp.name, m.name)


def test_dict_matches_reference(golden):
    for case in ("example", "edge"):
        d = os.path.join(golden, case)
        for sp in O.load_dict(os.path.join(d, "sequences.mf.dict")):
            st = O.build_stream(sp.files, base_dir=d)
            from blamm_b200 import synth
            counts = synth.counts_of(np.frombuffer(st.chars, dtype=np.uint8))
            assert counts == sp.counts and len(st.chars) == sp.tot_len
            assert st.seq_names == sp.seq_names


@pytest.mark.parametrize("case,exact", [("example", True), ("edge", False)])
def test_empirical_histograms_match_reference(golden, case, exact):
    """`blamm hist -e` (hist.cpp:70-160): oracle counts vs the files written by the compiled reference.  On the example
    (short motifs, default settings) every bin is identical; on `edge` the reference's sgemm edge kernels move a few
    scores by an ulp across a bin edge, so bins may trade single counts (total preserved)."""
    d = os.path.join(golden, case)
    motifs = O.load_jaspar(os.path.join(d, "motifs.jaspar"))
    for sp in O.load_dict(os.path.join(d, "sequences.mf.dict")):
        P, col_len = O.generate_matrix(motifs, sp.counts)
        mm = [O.max_min_score(np.ascontiguousarray(P[c, :4 * len(m)])) for c, m in enumerate(motifs)]
        mx = np.array([a for a, b in mm], np.float32); mn = np.array([b for a, b in mm], np.float32)
        st = O.build_stream(sp.files, 10_000_000, d)
        cnt = O.empirical_hist(st.chars, st.frag_start, P, col_len, mn, mx, 250)
        for c, m in enumerate(motifs):
            nb, hmn, hmx, ref = O.load_hist(os.path.join(d, "hist_e", "hist_%s_%s.dat" % (sp.name, m.name)))
            assert nb == 250 and int(ref.sum()) == int(cnt[c].sum())
            if exact:
                assert np.array_equal(ref, cnt[c]), (sp.name, m.name)
            else:
                assert int(np.abs(ref.astype(np.int64) - cnt[c].astype(np.int64)).sum()) <= 1e-4 * int(ref.sum()) + 4


REF_BIN = os.path.join(util.ROOT, "oracle", "_ref", "blamm")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/blamm (the compiled reference) has not been built")
@pytest.mark.parametrize("seed,flags,mode", [(77, ["-rc", "-pt", "0.001"], ("pt", 0.001, True)), (78, ["-rc"], ("rt", 0.95, True)),
                                             (79, ["-at", "7.5"], ("at", 7.5, False))])
def test_oracle_matches_live_reference_binary(tmp_path, seed, flags, mode):
    """Beyond the committed fixtures: the oracle's whole-`scan` restatement against the UNMODIFIED reference binary run here on
    fresh seeded inputs -- two manifest groups with distinct backgrounds, N runs, lower-case stretches, records shorter than a
    motif.  Motifs are at most 14 long, where the reference's sgemm sums in position order, so the sorted
    occurrence text must be identical byte for byte."""
    import subprocess
    from blamm_b200 import synth
    rng = np.random.default_rng(seed)
    synth.make_jaspar_like(str(tmp_path / "motifs.jaspar"), 30, seed, uniform_len=(5, 14))
    manifest = []
    for g, gc in enumerate((0.38, 0.5)):
        probs = ((1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2)
        seq = synth.random_acgt(260_000, seed * 10 + g, probs)
        for _ in range(5):
            a = int(rng.integers(0, len(seq) - 3000)); seq[a:a + int(rng.integers(1, 2000))] = ord("N")
            b = int(rng.integers(0, len(seq) - 3000)); seq[b:b + int(rng.integers(1, 2500))] |= 0x20
        recs = [("g%dchr1" % g, seq[:120_000]), ("g%dtiny" % g, seq[120_000:120_009]), ("g%dchr2 note" % g, seq[120_009:])]
        synth.write_fasta(str(tmp_path / ("g%d.fa" % g)), recs)
        manifest.append("grp%d\tg%d.fa\n" % (g, g))
    open(tmp_path / "seq.mf", "w").write("".join(manifest))
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    ob = os.path.join(util.ROOT, "oracle", "_ref", "openblas_dir.txt")
    if os.path.exists(ob):
        env["LD_LIBRARY_PATH"] = open(ob).read().strip() + ":" + env.get("LD_LIBRARY_PATH", "")
    for args in (["dict", "seq.mf"], ["hist", "motifs.jaspar", "seq.mf"], ["scan", "-t", "2"] + flags + ["motifs.jaspar", "seq.mf"]):
        r = subprocess.run([REF_BIN] + args, cwd=tmp_path, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    want = sorted(open(tmp_path / "occurrences.txt").read().splitlines(True))
    lines, det = O.scan("motifs.jaspar", "seq.mf", mode[0], mode[1], mode[2], histdir=".", base_dir=str(tmp_path))
    assert len(want) > 200 and sorted(lines) == want
