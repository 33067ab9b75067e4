#!/usr/bin/env bash
# Benchmark of the CLI's HOST pipeline without a GPU: `blamm-b200 scan` against the test-only stand-in library
# (tests/mock/mock_b200scan.cpp) in its synthetic-hit mode -- no scoring, a pseudo-random ordered hit list of configs[2]'s density
# (1.35e-4 hits per window and column) per chunk -- so that what is timed is the reader + packer, the workers' bookkeeping, the
# formatter and the writer at the rate a B200 would feed them.   usage: tools/host_pipeline_bench.sh [Mbp=400] [devices=1] [threads=nproc]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
MBP=${1:-400}; DEV=${2:-1}; THREADS=${3:-$(nproc)}
W=${WORKDIR:-/dev/shm/host_pipeline_bench}
mkdir -p $W $ROOT/tests/mock/_build
g++ -O2 -std=c++17 -shared -fPIC $ROOT/tests/mock/mock_b200scan.cpp -o $ROOT/tests/mock/_build/libb200scan.so -L$ROOT/oracle -loracle -Wl,-rpath,$ROOT/oracle -lpthread
if [ ! -f $W/seq.mf.dict ] || [ "$(cat $W/mbp 2>/dev/null)" != "$MBP" ]; then
  python - <<PY
import os, sys
sys.path.insert(0, "$ROOT")
from blamm_b200 import synth
W = "$W"
synth.make_jaspar_like(os.path.join(W, "motifs.jaspar"), 900, 2024)
per = 100_000_000
n = int($MBP * 1e6)
with open(os.path.join(W, "seq.mf"), "w") as mf:
    for g in range((n + per - 1) // per):
        seq = synth.random_acgt(min(per, n - g * per), 500 + g)
        q = len(seq) // 3
        synth.write_fasta(os.path.join(W, "g%02d.fa" % g), [("g%02d_chr%d" % (g, i + 1), seq[i * q:(i + 1) * q]) for i in range(3)])
        mf.write("group%02d\tg%02d.fa\n" % (g, g))
PY
  (cd $W && $ROOT/blamm_b200/lib/blamm-b200 dict seq.mf > /dev/null && $ROOT/blamm_b200/lib/blamm-b200 hist motifs.jaspar seq.mf > /dev/null)
  echo $MBP > $W/mbp
fi
cd $W
for pass in 1 2; do
  t0=$(date +%s.%N)
  LD_LIBRARY_PATH=$ROOT/tests/mock/_build:$LD_LIBRARY_PATH MOCK_B200SCAN_DEVICES=$DEV MOCK_B200SCAN_SYNTH_HITS=1.35e-4 BLAMM_B200_TIMING=1 \
    $ROOT/blamm_b200/lib/blamm-b200 scan -rc -pt 0.0001 -t $THREADS -o occ.txt motifs.jaspar seq.mf 2> timing.log | grep -E "Wrote|Using"
  python3 -c "import time,os; print('pass $pass: %.2f s wall, %.2f GB of text' % (time.time()-$t0, os.path.getsize('occ.txt')/1e9))"
  grep -E "reader|format|file write|turn|join|queue" timing.log
  rm -f occ.txt
done
