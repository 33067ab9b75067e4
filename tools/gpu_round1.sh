set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
nproc; lscpu | grep "Model name"
timeout 300 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "tc_debug rc=$?"
tail -40 gpurun_out/tc_debug.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench.log
timeout 600 python bench.py --steps 2 --warmup 3 --engine gather --no-cpu-baseline > gpurun_out/bench_gather.log 2>&1; echo "bench gather rc=$?"; tail -2 gpurun_out/bench_gather.log
