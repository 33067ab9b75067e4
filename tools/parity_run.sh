#!/usr/bin/env bash
# Near-threshold exception list at size: blamm-b200 and the reference's CPU BLAS path scan the same MBP of
# synthetic sequence with the 900-motif set (-rc -pt 1e-4, motifs up to 35 long); tools/parity_list.py compares the occurrence
# SETS and lists every difference with its distance from the threshold.  usage: tools/parity_run.sh [Mbp] [out prefix] [bias]
set -e
cd "$(dirname "$0")/.."
ROOT=$PWD; MBP=${1:-100}; OUT=${2:-$ROOT/gpurun_out/r2_parity}; BIAS=${3:-0}
W=$(mktemp -d); cd $W
python - <<PY
import sys; sys.path.insert(0, "$ROOT")
from blamm_b200 import synth
synth.make_jaspar_like("motifs.jaspar", 900, 2024)
n = int($MBP * 1e6); q = n // 4
seq = synth.random_acgt(n, 777, (0.295, 0.205, 0.205, 0.295) if $BIAS else (0.25, 0.25, 0.25, 0.25))
synth.write_fasta("genome.fa", [("chr%d" % (i + 1), seq[i * q:(i + 1) * q]) for i in range(4)])
open("genome.mf", "w").write("syn\tgenome.fa\n")
PY
B=$ROOT/blamm_b200/lib/blamm-b200; R=$ROOT/oracle/_ref/blamm
export OPENBLAS_NUM_THREADS=1
$R dict genome.mf > /dev/null; $R hist motifs.jaspar genome.mf > /dev/null
s=$(date +%s.%N); $B scan -rc -pt 0.0001 -o occ_b200.txt motifs.jaspar genome.mf > /dev/null; e=$(date +%s.%N)
echo "blamm-b200 scan: $(python -c "print('%.2f' % ($e - $s))") s, $(wc -l < occ_b200.txt) lines" | tee $OUT.log
s=$(date +%s.%N); $R scan -rc -pt 0.0001 -t $(nproc) -o occ_ref.txt motifs.jaspar genome.mf > /dev/null; e=$(date +%s.%N)
echo "reference scan -t $(nproc): $(python -c "print('%.2f' % ($e - $s))") s, $(wc -l < occ_ref.txt) lines" | tee -a $OUT.log
python $ROOT/tools/parity_list.py --ours occ_b200.txt --ref occ_ref.txt --motifs motifs.jaspar --manifest genome.mf --pt 0.0001 --rc --out ${OUT}_exceptions.txt | tee -a $OUT.log
cd /; rm -rf $W
