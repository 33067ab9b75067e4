#!/usr/bin/env bash
# Build compile-time variants of the filter kernel (local) or bench them (on the GPU box): tools/variants.sh build|run
set -e
cd "$(dirname "$0")/.."
V=blamm_b200/lib/variants
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared blamm_b200/csrc/b200scan.cu"
if [ -n "$VARSET" ]; then eval "declare -A VAR=( $VARSET )"; else
declare -A VAR=( [base]="" [ko_prod]="-DTC_KNOCKOUT=4" [ko_push]="-DTC_KNOCKOUT=8" [ko_prod_push]="-DTC_KNOCKOUT=12" [ko_mma_push]="-DTC_KNOCKOUT=10" )
fi
if [ "$1" = build ]; then
  mkdir -p $V
  for k in "${!VAR[@]}"; do $NV ${VAR[$k]} $EXTRA -o $V/$k.so & done; wait; ls -la $V
else
  mkdir -p gpurun_out
  for k in "${!VAR[@]}"; do
    [ -f $V/$k.so ] || continue
    echo "== $k"; B200_BENCH_DIAG=1 B200SCAN_LIB=$PWD/$V/$k.so timeout 300 python bench.py --mbp ${MBP:-50} --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ms/step %.2f  kernel_ms %.2f  value %.3e  frac %.3f  cand %d hits %d' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value'], d['roofline']['frac'], d['config']['candidates_per_step'], d['config']['hits_per_step']))
    elif 'rror' in l: print('   ', l.strip()[:200])
"
  done
fi
