#!/usr/bin/env bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bucket_order_kernel|rescore_tile_kernel' -s 4 -c 4 -f -o gpurun_out/r2_prof_small2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --packed > gpurun_out/r2_ncu_small2.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --no-cpu-baseline --steps 5 --packed 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('value %.3e ms/step %.2f' % (d['value'], d['ms_per_step']), d['e2e']['stages_ms_last_block'])
"
