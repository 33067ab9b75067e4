# second-session check on the GPU box: parity tests, bench line, CLI end to end (parallel ingest, early context creation)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader | head -1
nproc; lscpu | grep "Model name"
timeout 900 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench.log
timeout 400 bash tools/cli_e2e.sh 100 8 1 > gpurun_out/cli_e2e.log 2>&1; echo "cli_e2e rc=$?"; cat gpurun_out/cli_e2e.log
for t in 1 4 16; do BLAMM_B200_INGEST_THREADS=$t timeout 300 bash tools/ingest_bench.sh 400 >> gpurun_out/ingest.log 2>&1; done; cat gpurun_out/ingest.log
