#!/usr/bin/env python
"""Regenerates tests/golden/ by running the UNMODIFIED reference built in oracle/_ref (oracle/build_ref.sh).

Run in the build container only (needs /root/reference for the example inputs and oracle/_ref/{blamm,refdump}):
    python tests/golden/make_golden.py
Fixtures written (all small):
  example/          the reference's example inputs (data, copied verbatim) + every output of the README walk-through:
                    sequences.mf.dict, hist_*.dat, PWMthresholds.txt, occ_<mode>.txt (sorted) and refdump_<mode>.bin
  edge/             seeded synthetic inputs with the awkward FASTA cases (N runs, lower case, short records, CRLF,
                    blank lines, several files per group) + the same outputs
  synth2m/          2 Mbp x 40 motifs: refdump only (bit-level scores from the reference's sgemm path)
"""
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from blamm_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
ENV = dict(os.environ, OPENBLAS_NUM_THREADS="1")
MODES = {"pt_rc": (["-rc", "-pt", "0.0001"], ("pt", "0.0001", "1")),
         "pt_fwd": (["-pt", "0.0001"], ("pt", "0.0001", "0")),
         "rt_rc": (["-rc"], ("rt", "0.95", "1")),
         "at_rc": (["-rc", "-at", "9.5"], ("at", "9.5", "1"))}


def run(cmd, cwd):
    subprocess.run(cmd, cwd=cwd, env=ENV, check=True, stdout=subprocess.DEVNULL)


def reference_outputs(d, motifs, manifest, modes, t="1"):
    run([REF + "/blamm", "dict", manifest], d)
    run([REF + "/blamm", "hist", motifs, manifest], d)
    for gnu in [f for f in os.listdir(d) if f.endswith(".gnu")]:
        os.remove(os.path.join(d, gnu))
    for name, (cli, dump) in modes.items():
        run([REF + "/blamm", "scan", "-t", t] + cli + [motifs, manifest], d)
        lines = sorted(open(os.path.join(d, "occurrences.txt")).read().splitlines(True))
        open(os.path.join(d, "occ_%s.txt" % name), "w").write("".join(lines))
        if name == "pt_rc":
            shutil.copy(os.path.join(d, "PWMthresholds.txt"), os.path.join(d, "PWMthresholds_pt_rc.txt"))
        run([REF + "/refdump", dump[0], dump[1], dump[2], ".", motifs, manifest, "refdump_%s.bin" % name], d)
    os.remove(os.path.join(d, "occurrences.txt"))
    os.remove(os.path.join(d, "PWMthresholds.txt"))
    # empirical histograms (`blamm hist -e`, first 10,000,000 characters of every group)
    os.makedirs(os.path.join(d, "hist_e"), exist_ok=True)
    run([REF + "/blamm", "hist", "-e", "-t", t, "-H", "hist_e", motifs, manifest], d)
    for gnu in [f for f in os.listdir(os.path.join(d, "hist_e")) if f.endswith(".gnu")]:
        os.remove(os.path.join(d, "hist_e", gnu))


def make_example():
    d = os.path.join(HERE, "example")
    shutil.rmtree(d, ignore_errors=True)
    shutil.copytree("/root/reference/example", d)
    reference_outputs(d, "motifs.jaspar", "sequences.mf", MODES)


def make_edge():
    d = os.path.join(HERE, "edge")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(os.path.join(d, "seq"))
    rng = np.random.default_rng(77)
    pfms = synth.random_pfms([5, 6, 6, 8, 9, 11, 12, 12, 15, 16, 20, 23], rng)
    synth.write_jaspar(os.path.join(d, "motifs.jaspar"), pfms, prefix="ED")

    def seq(n, seed, probs=(0.3, 0.2, 0.2, 0.3)):
        return synth.random_acgt(n, seed, probs)

    # file b.fa sorts before a2.fa? no: std::set order is byte order -> "seq/a2.fa" < "seq/b.fa" < "seq/c.fa"
    s1 = seq(30000, 1); s1[1000:1700] = ord("N"); s1[5000] = ord("n"); s1[9000:9003] = ord("R")
    s1[12000:15000] = np.frombuffer(bytes(s1[12000:15000]).lower(), dtype=np.uint8)      # soft-masked stretch
    s2 = seq(7, 2)                                                                        # shorter than most motifs
    s3 = seq(20011, 3); s3[:40] = ord("N"); s3[-25:] = ord("N")
    s4 = seq(64, 4)
    synth.write_fasta(os.path.join(d, "seq", "a2.fa"), [("chrA desc ignored", s1), ("tiny", s2), ("chrB", s3)], width=60)
    with open(os.path.join(d, "seq", "b.fa"), "wb") as f:                                  # CRLF, blank lines, odd widths, no final newline
        f.write(b">crlf\tx\r\n" + bytes(seq(150, 5)) + b"\r\n\r\n" + bytes(seq(33, 6)) + b"\r\n>empty\n>afterempty\n\n" + bytes(s4))
    s5 = seq(50021, 7, (0.2, 0.3, 0.3, 0.2))
    synth.write_fasta(os.path.join(d, "seq", "c.fa"), [("g2chr1", s5)], width=71)
    open(os.path.join(d, "sequences.mf"), "w").write("grpA\tseq/b.fa\ngrpB\tseq/c.fa\ngrpA\tseq/a2.fa\n")
    open(os.path.join(d, "settings.cnf"), "w").write("MATRIX_S_W 37\nMATRIX_S_H 211\nMATRIX_P_TILE_MIN_ZERO_AREA 1\nFLUSHOUTPUT 10\n")
    modes = dict(MODES)
    modes["at_low"] = (["-rc", "-at", "6.5"], ("at", "6.5", "1"))
    reference_outputs(d, "motifs.jaspar", "sequences.mf", modes, t="3")


def make_synth2m():
    d = os.path.join(HERE, "synth2m")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    synth.make_jaspar_like(os.path.join(d, "motifs.jaspar"), 40, seed=1234)
    recs = [("chr%d" % (i + 1), synth.random_acgt(500000, 100 + i)) for i in range(4)]
    synth.write_fasta(os.path.join(d, "genome.fa"), recs)
    open(os.path.join(d, "sequences.mf"), "w").write("syn\tgenome.fa\n")
    run([REF + "/blamm", "dict", "sequences.mf"], d)
    run([REF + "/blamm", "hist", "motifs.jaspar", "sequences.mf"], d)
    run([REF + "/refdump", "pt", "0.0001", "1", ".", "motifs.jaspar", "sequences.mf", "refdump_pt_rc.bin"], d)
    # the inputs are regenerated from their seeds by the tests; keep only what the reference produced
    for f in os.listdir(d):
        if f.endswith(".gnu") or f == "genome.fa":
            os.remove(os.path.join(d, f))


if __name__ == "__main__":
    make_example()
    make_edge()
    make_synth2m()
    total = sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(HERE) for f in fs)
    print("golden fixtures: %.1f KiB" % (total / 1024))
