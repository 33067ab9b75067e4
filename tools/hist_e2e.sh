#!/usr/bin/env bash
# `hist -e` (empirical score histograms) end to end on the GPU box, next to the reference binary on the same inputs:
# byte-compares every hist_*.dat and prints both wall times.  usage: tools/hist_e2e.sh [Mbp] [motifs]
set -e
cd "$(dirname "$0")/.."
ROOT=$PWD; MBP=${1:-10}; NM=${2:-200}
W=$(mktemp -d); cd $W
python - <<PY
import sys; sys.path.insert(0, "$ROOT")
from blamm_b200 import synth
synth.make_jaspar_like("motifs.jaspar", $NM, 2024)
n = int($MBP * 1e6); q = n // 4
seq = synth.random_acgt(n, 4242)
synth.write_fasta("genome.fa", [("chr%d" % (i + 1), seq[i * q:(i + 1) * q]) for i in range(4)])
open("genome.mf", "w").write("syn\tgenome.fa\n")
PY
B=$ROOT/blamm_b200/lib/blamm-b200; R=$ROOT/oracle/_ref/blamm
export OPENBLAS_NUM_THREADS=1
t() { local s=$(date +%s.%N); "$@" > log.txt 2>&1 || { cat log.txt; exit 1; }; python -c "print('%.2f' % ($(date +%s.%N) - $s))"; }
$B dict genome.mf > /dev/null
mkdir -p hb hr
L=$(python -c "print(int($MBP * 1e6))")
echo "hist -e  b200      ($MBP Mbp x $NM motifs): $(t $B hist -e -l $L -H hb motifs.jaspar genome.mf) s"
echo "hist -e  reference ($MBP Mbp x $NM motifs, -t $(nproc)): $(t $R hist -e -l $L -t $(nproc) -H hr motifs.jaspar genome.mf) s"
python - <<PY
# byte compare; for files that differ, how far: the reference scores through sgemm (re-associated sums for longer motifs), so
# a score within an ulp of a bin edge can land in the neighbouring bin
import os
n = bad = moved = 0; worst = 0; tot = 0
for f in sorted(os.listdir("hr")):
    if not f.endswith(".dat"): continue
    n += 1
    a, b = open("hr/" + f).read(), open("hb/" + f).read()
    if a == b: continue
    bad += 1
    ra = [l.split() for l in a.splitlines()]; rb = [l.split() for l in b.splitlines()]
    assert ra[0] == rb[0], (f, ra[0], rb[0]); ra, rb = ra[1:], rb[1:]
    assert len(ra) == len(rb), f
    for x, y in zip(ra, rb):
        assert x[:-1] == y[:-1], (f, x, y)                # same bin positions, only counts may move
        d = abs(int(float(x[-1])) - int(float(y[-1]))); moved += d; worst = max(worst, d)
    tot += sum(int(float(x[-1])) for x in ra)
print("histogram files: %d, byte-identical: %d, differing: %d (observations that changed bin: %d of %d, largest per-bin difference %d)" % (n, n - bad, bad, moved // 2, tot, worst))
PY
cd /; rm -rf $W
