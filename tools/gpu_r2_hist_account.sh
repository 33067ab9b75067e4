#!/usr/bin/env bash
# Phase account of `hist -e` on one GPU: 24 groups x 8 Mbp x 900 motifs (BLAMM_B200_TIMING=1), two passes.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
W=/dev/shm/hist_acc; rm -rf $W; mkdir -p $W
timeout 100 python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
from blamm_b200 import synth
W = "/dev/shm/hist_acc"
t0 = time.time()
synth.make_jaspar_like(os.path.join(W, "motifs.jaspar"), 900, 2024)
with open(os.path.join(W, "sequences.mf"), "w") as mf:
    for g in range(24):
        gc = 0.36 + 0.12 * g / 23
        seq = synth.random_acgt(8_000_000, 500 + g, (0.5 - gc / 2, gc / 2, gc / 2, 0.5 - gc / 2))
        synth.write_fasta(os.path.join(W, "g%02d.fa" % g), [("g%02d_chr1" % g, seq)])
        mf.write("group%02d\tg%02d.fa\n" % (g, g))
print("inputs: %.1f s" % (time.time() - t0))
PY
cd $W
CLI=$GRAFT_REPO_ROOT/blamm_b200/lib/blamm-b200
$CLI dict sequences.mf > /dev/null
for pass in 1 2; do
  t0=$(date +%s.%N)
  BLAMM_B200_TIMING=1 timeout 60 $CLI hist -e -l 9000000 -g 1 motifs.jaspar sequences.mf 2> $GRAFT_REPO_ROOT/gpurun_out/r2_hist_account_$pass.log > /dev/null
  echo "hist -e pass $pass: rc=$? $(echo "$(date +%s.%N) - $t0" | bc -l 2>/dev/null || python3 -c "import time;print(time.time()-$t0)") s wall" | tee -a $GRAFT_REPO_ROOT/gpurun_out/r2_hist_account_$pass.log
  cat $GRAFT_REPO_ROOT/gpurun_out/r2_hist_account_$pass.log
done
ls | grep -c "^hist_"
rm -rf $W
