#!/usr/bin/env bash
# Closing call of round 2: hist -e kernel 3 (lanes count different columns at a time) next to kernel 2 on 24 groups x 8 Mbp x
# 900 motifs (files must be identical), then the whole GPU parity suite on the final build (default hist kernel = 3).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
W=/dev/shm/hist_acc; rm -rf $W; mkdir -p $W
timeout 60 python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from blamm_b200 import synth
W = "/dev/shm/hist_acc"
synth.make_jaspar_like(os.path.join(W, "motifs.jaspar"), 900, 2024)
with open(os.path.join(W, "sequences.mf"), "w") as mf:
    for g in range(24):
        gc = 0.36 + 0.12 * g / 23
        seq = synth.random_acgt(8_000_000, 500 + g, (0.5 - gc / 2, gc / 2, gc / 2, 0.5 - gc / 2))
        if g % 3 == 0:
            seq[1_000_000:1_400_000] |= 0x20
            seq[3_000_000:3_000_700] = ord("N")
        synth.write_fasta(os.path.join(W, "g%02d.fa" % g), [("g%02d_chr1" % g, seq)])
        mf.write("group%02d\tg%02d.fa\n" % (g, g))
PY
cd $W
CLI=$GRAFT_REPO_ROOT/blamm_b200/lib/blamm-b200
$CLI dict sequences.mf > /dev/null
for v in 2 3 3; do
  mkdir -p h$v
  t0=$(date +%s.%N)
  B200SCAN_HIST_KERNEL=$v BLAMM_B200_TIMING=1 timeout 40 $CLI hist -e -l 9000000 -g 1 -H h$v motifs.jaspar sequences.mf 2> $GRAFT_REPO_ROOT/gpurun_out/r2_hist_kernel_$v.log > /dev/null
  echo "kernel $v: rc=$? $(python3 -c "import time;print('%.2f' % (time.time()-$t0))") s wall" | tee -a $GRAFT_REPO_ROOT/gpurun_out/r2_hist_kernel_$v.log
  grep -E "hist_block|last kernels" $GRAFT_REPO_ROOT/gpurun_out/r2_hist_kernel_$v.log
done
if diff -rq h2 h3 > /dev/null; then echo "kernel 3: all $(ls h3 | wc -l) files identical to kernel 2"; else echo "kernel 3: FILES DIFFER"; diff -rq h2 h3 | head -5; fi
rm -rf $W
cd $GRAFT_REPO_ROOT
timeout 105 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest_gpu.log
