#!/usr/bin/env python
"""Near-threshold exception lister (BASELINE.json north_star: "identical occurrence set ... scores within 1e-4 absolute; any
discrepancy is allowed only for hits whose reference score lies within that tolerance of the threshold, and such hits are
listed").

Compares two occurrence files of the same run -- blamm-b200's and the reference's (CPU BLAS path, pwmscan.cpp:104-133) -- as
SETS of (sequence, motif, start, end, strand), whatever order the lines come in, and

  * lists every occurrence present in only one file with its score, the threshold of its (group, motif) and |score - thr|;
  * fails (exit 1) if such an occurrence lies further than --tol (1e-4) from its threshold, allowing for the 6 significant
    digits the score column is printed with (`%g`: half a unit of the last printed digit);
  * reports the largest |score difference| over the common occurrences and how many exceed --tol.

Thresholds are recomputed with the host model (libblammhost.so: P and thresholds bit-identical to the reference's, see
tests/test_gpu_parity.py::test_matrix_and_thresholds_match_reference) from the same motif file, manifest .dict and histograms.
Sorting is done by `sort` (LC_ALL=C), the merge streams both files: 2.5e7 lines per side take about two minutes.

usage: parity_list.py --ours occ_b200.txt --ref occ_ref.txt --motifs motifs.jaspar --manifest sequences.mf
                      [--histdir DIR] (--pt P | --at A | --rt R) [--rc] [--tol 1e-4] [--out list.txt]
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def keyed_sorted(path: str, out: str) -> None:
    """GTF-like line (seq, 'blamm', motif, start, end, score, strand, '.', '.') -> 'seq\tmotif\tstart\tend\tstrand\tscore', sorted bytewise."""
    cmd = ("awk -F'\\t' 'BEGIN{OFS=\"\\t\"}{print $1,$3,$4,$5,$7,$6}' %s | LC_ALL=C sort -S 2G --parallel=%d -o %s"
           % (path, min(16, os.cpu_count() or 1), out))
    subprocess.run(cmd, shell=True, check=True, executable="/bin/bash")


def print_resolution(score: float) -> float:
    """Half a unit of the 6th significant digit of `score` (what `%g` keeps)."""
    import math
    a = abs(score)
    if a == 0.0:
        return 5e-7
    return 0.5 * 10.0 ** (math.floor(math.log10(a)) - 5)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ours", required=True); ap.add_argument("--ref", required=True)
    ap.add_argument("--motifs", required=True); ap.add_argument("--manifest", required=True)
    ap.add_argument("--histdir", default="")
    g = ap.add_mutually_exclusive_group(required=True)
    g.add_argument("--pt", type=float); g.add_argument("--at", type=float); g.add_argument("--rt", type=float)
    ap.add_argument("--rc", action="store_true")
    ap.add_argument("--tol", type=float, default=1e-4)
    ap.add_argument("--out", default="")
    ap.add_argument("--max-list", type=int, default=100000)
    args = ap.parse_args()

    from blamm_b200 import capi
    from oracle import oracle as O                                   # .dict reader only (test infrastructure; this is a checker)
    mode, value = ("pt", args.pt) if args.pt is not None else ("at", args.at) if args.at is not None else ("rt", args.rt)
    ms = capi.MotifSet(args.motifs, revcompl=args.rc)
    histdir = args.histdir or os.path.dirname(os.path.abspath(args.manifest))
    thr_of = {}                                                       # (sequence name) -> {(motif name, strand): threshold}
    for sp in O.load_dict(args.manifest + ".dict"):
        P, col_len, is_rc = ms.generate_matrix(sp.counts)
        thr = ms.thresholds(mode, value, sp.name, histdir)
        table = {(ms.names[c].encode(), b"-" if is_rc[c] else b"+"): float(thr[c]) for c in range(ms.n_cols)}
        for name in sp.seq_names:
            thr_of[name.encode()] = table

    tmp = tempfile.mkdtemp(prefix="parity_")
    a_path, b_path = os.path.join(tmp, "ours.k"), os.path.join(tmp, "ref.k")
    keyed_sorted(args.ours, a_path); keyed_sorted(args.ref, b_path)

    only = []                                                         # (which, key, score, thr, dist)
    n_common = n_a = n_b = n_over = 0
    max_diff, max_rel, worst = 0.0, 0.0, None

    def rows(path):
        with open(path, "rb") as f:
            for line in f:
                k, _, s = line.rstrip(b"\n").rpartition(b"\t")
                yield k, s

    def note(which, k, s):
        f = k.split(b"\t")
        score = float(s)
        thr = thr_of[f[0]][(f[1], f[4])]
        only.append((which, k.decode(), score, thr, abs(score - thr)))

    ia, ib = rows(a_path), rows(b_path)
    ka = next(ia, None); kb = next(ib, None)
    while ka is not None or kb is not None:
        if kb is None or (ka is not None and ka[0] < kb[0]):
            n_a += 1; note("only-b200", *ka); ka = next(ia, None)
        elif ka is None or kb[0] < ka[0]:
            n_b += 1; note("only-ref ", *kb); kb = next(ib, None)
        else:
            n_a += 1; n_b += 1; n_common += 1
            if ka[1] != kb[1]:
                x, y = float(ka[1]), float(kb[1])
                d = abs(x - y)
                if d > max_diff:
                    max_diff, worst = d, (ka[0].decode(), x, y)
                max_rel = max(max_rel, d / max(1.0, abs(y)))
                if d > args.tol + print_resolution(x) + print_resolution(y):
                    n_over += 1
            ka = next(ia, None); kb = next(ib, None)

    bad = [o for o in only if o[4] > args.tol + print_resolution(o[2])]
    lines = ["# parity_list.py: ours = %s (%d occurrences), reference = %s (%d occurrences), %s %g%s, tolerance %g"
             % (args.ours, n_a, args.ref, n_b, mode, value, " -rc" if args.rc else "", args.tol),
             "# common occurrences: %d; largest |score difference| %.3g (%.3g relative to max(1, |s|))%s; differences above the tolerance "
             "(+ print resolution): %d" % (n_common, max_diff, max_rel, (" at %s: %g vs %g" % worst) if worst else "", n_over),
             "# occurrences in one file only: %d (b200 only %d, reference only %d); of these further than the tolerance from their threshold: %d"
             % (len(only), sum(o[0] == "only-b200" for o in only), sum(o[0] != "only-b200" for o in only), len(bad)),
             "# which\tsequence\tmotif\tstart\tend\tstrand\tscore\tthreshold\t|score-threshold|"]
    for o in only[: args.max_list]:
        lines.append("%s\t%s\t%.7g\t%.9g\t%.3g" % o)
    text = "\n".join(lines) + "\n"
    sys.stdout.write(text if len(only) <= 50 else "\n".join(lines[:54]) + "\n...\n")
    if args.out:
        open(args.out, "w").write(text)
    for p in (a_path, b_path):
        os.remove(p)
    os.rmdir(tmp)
    ok = not bad and n_over == 0
    print("PARITY %s" % ("OK: identical occurrence sets" if ok and not only else
                          "OK: every set difference lies within the tolerance of its threshold (listed)" if ok else "FAILED"))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
