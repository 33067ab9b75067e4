#!/usr/bin/env bash
# Round 2, final validation of the committed tree: smoke, GPU parity suite, bench (both arms' GPU side), CLI end to end.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "bench rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_c2.json'):
    if l.startswith('{'):
        d = json.loads(l); print('c2: ms/step %.3f value %.3e e2e %.3e (%.2f ms) frac %.3f pipe %.3f launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['pipe']['frac'], d['gpu_launches']), d['cpu_baseline']['value'], d['cpu_baseline'].get('gpu_reference'))
PY
tail -n 3 gpurun_out/r2_bench_c2.err
BLAMM_B200_TIMING=1 timeout 600 bash tools/cli_e2e.sh 100 8 1 > gpurun_out/r2_cli_e2e.log 2>&1; grep -E "scan  |identical|lines" gpurun_out/r2_cli_e2e.log
