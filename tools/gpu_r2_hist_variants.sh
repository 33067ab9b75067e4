#!/usr/bin/env bash
# hist -e kernels side by side on one GPU (B200SCAN_HIST_KERNEL = 0: round-1 kernel, 1: + aggregated bin counts, 2: position-row
# weights + aggregated counts): the GPU histogram tests under each, then 24 groups x 8 Mbp x 900 motifs with a phase account
# (BLAMM_B200_TIMING=1); the .dat files of the three kernels must be identical.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 0 1 2; do
  B200SCAN_HIST_KERNEL=$v timeout 60 python -m pytest tests -x -q -m gpu -k "empirical" 2>&1 | tail -1 | sed "s/^/kernel $v tests: /"
done
W=/dev/shm/hist_acc; rm -rf $W; mkdir -p $W
timeout 100 python - <<'PY'
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
            seq[1_000_000:1_400_000] |= 0x20          # a soft-masked stretch: the masked kernel instance
            seq[3_000_000:3_000_700] = ord("N")       # and a gap: fragments
        synth.write_fasta(os.path.join(W, "g%02d.fa" % g), [("g%02d_chr1" % g, seq)])
        mf.write("group%02d\tg%02d.fa\n" % (g, g))
PY
cd $W
CLI=$GRAFT_REPO_ROOT/blamm_b200/lib/blamm-b200
$CLI dict sequences.mf > /dev/null
for v in 0 1 2 2; do
  mkdir -p h$v
  t0=$(date +%s.%N)
  B200SCAN_HIST_KERNEL=$v BLAMM_B200_TIMING=1 timeout 60 $CLI hist -e -l 9000000 -g 1 -H h$v motifs.jaspar sequences.mf 2> $GRAFT_REPO_ROOT/gpurun_out/r2_hist_kernel_$v.log > /dev/null
  echo "kernel $v: rc=$? $(python3 -c "import time;print('%.2f' % (time.time()-$t0))") s wall" | tee -a $GRAFT_REPO_ROOT/gpurun_out/r2_hist_kernel_$v.log
  grep -E "hist_block|last kernels" $GRAFT_REPO_ROOT/gpurun_out/r2_hist_kernel_$v.log
done
for v in 1 2; do
  if diff -rq h0 h$v > /dev/null; then echo "kernel $v: all $(ls h$v | wc -l) files identical to kernel 0"; else echo "kernel $v: FILES DIFFER"; diff -rq h0 h$v | head -5; fi
done
rm -rf $W
