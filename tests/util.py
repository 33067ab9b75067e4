"""Shared helpers for the test-suite (tests may use the oracle; the product never does)."""
import os
import shutil

import numpy as np

from blamm_b200 import synth
from oracle import oracle as O
from oracle.refdump_io import read_refdump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MODES = {"pt_rc": ("pt", 1e-4, True), "pt_fwd": ("pt", 1e-4, False), "rt_rc": ("rt", 0.95, True),
         "at_rc": ("at", 9.5, True), "at_low": ("at", 6.5, True)}


def materialise_synth2m(dst):
    """tests/golden/synth2m keeps only what the reference produced; the inputs come back from their seeds."""
    shutil.copytree(os.path.join(GOLDEN, "synth2m"), dst, dirs_exist_ok=True)
    synth.make_jaspar_like(os.path.join(dst, "motifs.jaspar"), 40, seed=1234)
    recs = [("chr%d" % (i + 1), synth.random_acgt(500000, 100 + i)) for i in range(4)]
    synth.write_fasta(os.path.join(dst, "genome.fa"), recs)
    return dst


def oracle_case(case_dir, mode_key, lower_fold=False):
    mode, value, rc = MODES[mode_key]
    return O.scan("motifs.jaspar", "sequences.mf", mode, value, rc, histdir=".", base_dir=case_dir, lower_fold=lower_fold)


def hit_keys(seq, pos, col, score):
    """Canonical, sortable identity of hits including the score bits."""
    return sorted(zip(np.asarray(seq).tolist(), np.asarray(pos).tolist(), np.asarray(col).tolist(),
                      np.asarray(score, dtype=np.float32).view(np.uint32).tolist()))


def refdump_keys(species_dump, field="score"):
    h = species_dump["hits"]
    return hit_keys(h["seq"], h["pos"], h["col"], h[field])


def compare_with_refdump(d, r, blas_bit_exact):
    """Oracle result `d` (one species of oracle.scan) against the reference dump `r` of the same species:
       * identical occurrence set (record, position, column) -- always;
       * oracle score == the reference's naive-path score (Motif::getScore, in-order float adds) BIT FOR BIT for
         every window without lower-case characters (the naive path folds case, the BLAS path zeroes it);
       * oracle score vs the reference's BLAS-path score: bit-identical where `blas_bit_exact` (short motifs,
         default settings), else within 1e-5 -- OpenBLAS re-associates the K loop for long motifs / edge
         tiles, so the reference's own BLAS and naive paths differ by a few ulp there (north_star allows 1e-4)."""
    from oracle import oracle as O
    ref_col = {(c["name"], c["rc"]): j for j, c in enumerate(r["cols"])}
    cmap = np.array([ref_col[(m.name, m.revcomp)] for m in d["motifs"]])
    seq, spos = O.stream_to_seq(d["stream"], d["pos"])
    order = np.lexsort((cmap[d["col"]], spos, seq))
    h = r["hits"][np.lexsort((r["hits"]["col"], r["hits"]["pos"], r["hits"]["seq"]))]
    assert len(order) == len(h)
    assert np.array_equal(seq[order], h["seq"]) and np.array_equal(spos[order], h["pos"]) and np.array_equal(cmap[d["col"]][order], h["col"])
    mine = d["score"][order]
    chars = d["stream"].chars
    upper = np.array([chars[int(p):int(p) + int(d["col_len"][c])].isupper() for p, c in zip(d["pos"][order], d["col"][order])], dtype=bool)
    assert np.array_equal(mine[upper].view(np.uint32), h["naive"][upper].view(np.uint32))
    if blas_bit_exact:
        assert np.array_equal(mine.view(np.uint32), h["score"].view(np.uint32))
    else:
        assert np.max(np.abs(mine - h["score"]), initial=0.0) <= 1e-5
    return cmap


def random_case(seed, n_motifs=24, n_nt=200_000, len_range=(5, 30), with_gaps=True, lower=False):
    """Seeded in-memory case: (P, col_len, thr, chars, frag_start) with -rc columns and -at/-rt style thresholds."""
    rng = np.random.default_rng(seed)
    lengths = rng.integers(len_range[0], len_range[1] + 1, size=n_motifs)
    pfms = synth.random_pfms(lengths, rng)
    motifs = O.add_revcompl(sorted([O.Motif("M%03d" % i, p.tolist()) for i, p in enumerate(pfms)], key=len))
    probs = rng.dirichlet(np.full(4, 20.0))
    chars = synth.random_acgt(n_nt, seed + 1, probs)
    P, col_len = O.generate_matrix(motifs, synth.counts_of(chars))
    thr = np.zeros(len(motifs), dtype=np.float32)
    for c, m in enumerate(motifs):
        mx, mn = O.max_min_score(np.ascontiguousarray(P[c, :4 * len(m)]))
        thr[c] = np.float32(min(float(mx) * 0.55, 9.0 + 0.2 * len(m)))
    frag = [0]
    if with_gaps:
        k = int(rng.integers(3, 40))
        frag = sorted(set([0] + rng.integers(1, n_nt, size=k).tolist() + [n_nt - 3, 17, 18]))
    if lower:
        a = int(rng.integers(0, n_nt // 2)); b = a + int(rng.integers(10, n_nt // 4))
        chars[a:b] = np.frombuffer(bytes(chars[a:b]).lower(), dtype=np.uint8)
    return dict(P=P, col_len=col_len, thr=thr, chars=chars, frag_start=np.array(frag, dtype=np.uint64), motifs=motifs)
