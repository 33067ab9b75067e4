#!/usr/bin/env bash
# End-to-end CLI run on the GPU box: FASTA in -> occurrences.txt out, next to the reference binary on a subset
# (same inputs, sorted outputs compared byte for byte).  usage: [SOFTMASK=0.5] tools/cli_e2e.sh [Mbp] [subset Mbp] [gpus]
# SOFTMASK = fraction of the sequence written in lower case (runs of 1..3000), as in repeat-masked genomes
set -e
cd "$(dirname "$0")/.."
ROOT=$PWD; MBP=${1:-100}; SUB=${2:-8}; GPUS=${3:-1}
W=$(mktemp -d); cd $W
python - <<PY
import sys; sys.path.insert(0, "$ROOT")
from blamm_b200 import synth
synth.make_jaspar_like("motifs.jaspar", 900, 2024)
n = int($MBP * 1e6); q = n // 4
seq = synth.random_acgt(n, 4242)
soft = float("${SOFTMASK:-0}")
if soft > 0:
    import numpy as np
    rng = np.random.default_rng(9); p = 0
    while p < n:
        run = int(rng.integers(1, 3000))
        if rng.random() < soft: seq[p:p + run] |= 0x20
        p += run
synth.write_fasta("genome.fa", [("chr%d" % (i + 1), seq[i * q:(i + 1) * q]) for i in range(4)])
m = int($SUB * 1e6) // 4
synth.write_fasta("subset.fa", [("chr%d" % (i + 1), seq[i * q:i * q + m]) for i in range(4)])
open("genome.mf", "w").write("syn\tgenome.fa\n"); open("subset.mf", "w").write("syn\tsubset.fa\n")
PY
B=$ROOT/blamm_b200/lib/blamm-b200; R=$ROOT/oracle/_ref/blamm
export OPENBLAS_NUM_THREADS=1 BLAMM_B200_TIMING=1
t() { local s=$(date +%s.%N); "$@" > log.txt 2>&1 || { cat log.txt; exit 1; }; python -c "print('%.2f' % ($(date +%s.%N) - $s))"; }
echo "dict   b200: $(t $B dict genome.mf) s"
echo "hist   b200: $(t $B hist motifs.jaspar genome.mf) s"
echo "scan   b200 ($MBP Mbp x 1800 cols, -rc -pt 1e-4, $GPUS GPU): $(t $B scan -rc -pt 0.0001 -g $GPUS -o occ_full.txt motifs.jaspar genome.mf) s; $(wc -l < occ_full.txt) lines, $(du -m occ_full.txt | cut -f1) MB"
grep timing log.txt; tail -2 log.txt
$B dict subset.mf > /dev/null; cp genome.mf.dict /dev/null
# the subset has its own background -> its own histograms / thresholds; both programs read the same files
$R dict subset.mf > /dev/null; $R hist motifs.jaspar subset.mf > /dev/null
echo "scan   b200 subset ($SUB Mbp): $(t $B scan -rc -pt 0.0001 -o occ_b200.txt motifs.jaspar subset.mf) s"
echo "scan   reference subset ($SUB Mbp, -t $(nproc)): $(t $R scan -rc -pt 0.0001 -t $(nproc) -o occ_ref.txt motifs.jaspar subset.mf) s"
echo "lines: b200 $(wc -l < occ_b200.txt) reference $(wc -l < occ_ref.txt)"
python - <<PY
# identical occurrence set (sequence, start, end, strand, motif); scores within 1e-4 (the reference's BLAS path
# re-associates long dot products, so its printed 6th digit can differ from the in-order sum)
def load(f):
    d = {}
    for l in open(f):
        c = l.rstrip("\n").split("\t")
        d[(c[0], c[2], c[3], c[4], c[6])] = float(c[5])
    return d
a, b = load("occ_b200.txt"), load("occ_ref.txt")
print("occurrence sets identical:", set(a) == set(b), " only-b200", len(set(a) - set(b)), " only-ref", len(set(b) - set(a)))
common = set(a) & set(b)
print("max |score diff| over %d common hits: %.3g" % (len(common), max(abs(a[k] - b[k]) / max(1.0, abs(b[k])) for k in common)), "(relative to max(1,|s|): 6 printed digits)")
import subprocess
same = subprocess.run("cmp -s <(LC_ALL=C sort occ_b200.txt) <(LC_ALL=C sort occ_ref.txt)", shell=True, executable="/bin/bash").returncode == 0
print("sorted text byte-identical:", same)
PY
cd /; rm -rf $W
