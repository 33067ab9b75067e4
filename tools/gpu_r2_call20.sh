#!/usr/bin/env bash
# work-item size of the filter (TC_SPAN) variants + CLI end to end with the recycled text buffers
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in span32k span128k; do
  echo "== $v"; B200SCAN_LIB=$PWD/blamm_b200/lib/variants/$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --packed 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ms/step %.3f  kernel_ms %.3f  value %.3e' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value']))
    elif 'rror' in l: print('   ', l.strip()[:200])
"
done
echo "== default"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --packed 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ms/step %.3f  kernel_ms %.3f  value %.3e' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value']))
"
BLAMM_B200_TIMING=1 timeout 600 bash tools/cli_e2e.sh 100 8 1 > gpurun_out/r2_cli_e2e.log 2>&1; grep -E "scan  |timing|identical|lines" gpurun_out/r2_cli_e2e.log
