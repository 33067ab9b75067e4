#!/usr/bin/env bash
# Round 2, call 5: parity suite + bench with the two-level ordering kernels, launch list, CLI end to end + start-up split,
# near-threshold exception list at 100 Mbp.   gpurun --timeout 1500 -- 'bash tools/gpu_r2_call5.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/r2_bench_c2.json; tail -5 gpurun_out/r2_bench_c2.err
timeout 300 python bench.py --no-cpu-baseline --hits 12 --steps 5 > gpurun_out/r2_bench_c2_hits12.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
BLAMM_B200_TIMING=1 timeout 600 bash tools/cli_e2e.sh 100 8 1 > gpurun_out/r2_cli_e2e.log 2>&1; tail -32 gpurun_out/r2_cli_e2e.log
timeout 300 bash tools/cli_startup.sh > gpurun_out/r2_cli_startup.log 2>&1; tail -30 gpurun_out/r2_cli_startup.log
timeout 900 bash tools/parity_run.sh 100 $PWD/gpurun_out/r2_parity 2>&1 | tail -30
