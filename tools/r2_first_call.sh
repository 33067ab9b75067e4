#!/usr/bin/env bash
# First GPU call of the next round (prepared at the end of round 1, when the GPU budget was spent).
#   here:        VARSET='[base]="" [span128k]="-DTC_SPAN=131072"' tools/variants.sh build
#   on the box:  gpurun --timeout 400 -- 'bash tools/r2_first_call.sh'
#   then, 8 GPUs (charged 8x): gpurun --gpus 8 --timeout 200 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8
#       --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8_hits12.log 2>&1'
# Questions it answers: (1) work items of 131,072 windows (225 KB of shared memory) against 65,536; (2) e2e with K = 5 / 10 / 20
# steps (how much of the 1.5 ms over the device-resident step is pipeline fill and drain: DESIGN.md section 8); (3) the CLI with
# the radix-sort / fast-%g writer (the 0.36 s "sort + format" phase of profiles/r01_cli_e2e.log).
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
VARSET='[base]="" [span128k]="-DTC_SPAN=131072"' MBP=100 bash tools/variants.sh run > gpurun_out/r2_variants.log 2>&1; cat gpurun_out/r2_variants.log
for k in 5 10 20; do
  timeout 200 python bench.py --steps $k --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('steps $k: value %.3e (%.2f ms)  e2e %.3e (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
" | tee -a gpurun_out/r2_e2e_steps.log
done
bash tools/cli_e2e.sh 100 8 1 > gpurun_out/r2_cli_e2e.log 2>&1; grep -E "scan   b200|timing|identical" gpurun_out/r2_cli_e2e.log
