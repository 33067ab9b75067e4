#!/usr/bin/env bash
# Round 2, call 6: parity suite on the pipelined CLI, bench (staged order kernel), configs c3 / c5 at a tenth of their size and c4 at
# 300 Mbp on one GPU (the full sizes run on 8 GPUs: tools/gpu_r2_scale8.sh).   gpurun --timeout 1500 -- 'bash tools/gpu_r2_call6.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "bench rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_c2.json'):
    if l.startswith('{'):
        d = json.loads(l); print('c2: ms/step %.3f value %.3e e2e %.3e (%.2f ms) frac %.3f pipe %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['pipe']['frac']), d['e2e']['stages_ms_last_block'])
PY
tail -3 gpurun_out/r2_bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
BLAMM_B200_TIMING=1 timeout 600 bash tools/cli_e2e.sh 100 8 1 > gpurun_out/r2_cli_e2e.log 2>&1; grep -E "scan  |timing|identical|lines" gpurun_out/r2_cli_e2e.log
timeout 900 python bench.py --config c3 --gbp 0.31 --steps 1 --warmup 1 > gpurun_out/r2_bench_c3_small.json 2> gpurun_out/r2_bench_c3_small.err; echo "c3 rc=$?"; tail -c 2500 gpurun_out/r2_bench_c3_small.json; tail -5 gpurun_out/r2_bench_c3_small.err
timeout 900 python bench.py --config c5 --gbp 0.31 --steps 1 --warmup 0 > gpurun_out/r2_bench_c5_small.json 2> gpurun_out/r2_bench_c5_small.err; echo "c5 rc=$?"; tail -c 1500 gpurun_out/r2_bench_c5_small.json; tail -5 gpurun_out/r2_bench_c5_small.err
timeout 900 python bench.py --config c4 --mbp 300 --steps 2 --warmup 3 > gpurun_out/r2_bench_c4_small.json 2> gpurun_out/r2_bench_c4_small.err; echo "c4 rc=$?"; tail -c 3000 gpurun_out/r2_bench_c4_small.json; tail -5 gpurun_out/r2_bench_c4_small.err
