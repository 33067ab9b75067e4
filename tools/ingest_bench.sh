#!/usr/bin/env bash
# FASTA ingest throughput of `blamm-b200 dict` (parse + filter + nucleotide counts; no GPU involved).
# usage: [BLAMM_B200_INGEST_THREADS=n] tools/ingest_bench.sh [Mbp]
set -e
cd "$(dirname "$0")/.."
ROOT=$PWD; MBP=${1:-400}
W=${INGEST_DIR:-/tmp/ingest_bench_$MBP}
if [ ! -f $W/g.fa ]; then
mkdir -p $W
python - <<PY
import sys; sys.path.insert(0, "$ROOT")
from blamm_b200 import synth
n = int($MBP * 1e6); q = n // 4
seq = synth.random_acgt(n, 3)
for a in range(q // 3, n, 37_000_000): seq[a:a + 50_000] = ord("N")
synth.write_fasta("$W/g.fa", [("chr%d" % (i + 1), seq[i * q:(i + 1) * q]) for i in range(4)])
open("$W/s.mf", "w").write("syn\t$W/g.fa\n")
PY
fi
cd $W
cat g.fa > /dev/null
best=999
for rep in 1 2 3 4 5; do
  s=$(date +%s.%N); $ROOT/blamm_b200/lib/blamm-b200 dict s.mf > /dev/null; e=$(date +%s.%N)
  best=$(python -c "print(min($best, $e - $s))")
done
python -c "import os; b = os.path.getsize('g.fa'); print('ingest threads=%s: %.3f s best of 5 for %.0f MB -> %.2f GB/s' % (os.environ.get('BLAMM_B200_INGEST_THREADS', 'all'), $best, b / 1e6, b / 1e9 / $best))"
