# confirmation run of the 12-byte hit records: parity tests, bench line with 12- and 16-byte records, CLI next to the reference
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout 100 python bench.py --steps 5 --warmup 3 --hits 16 --no-cpu-baseline > gpurun_out/bench_hits16.log 2>&1; echo "bench16 rc=$?"; tail -1 gpurun_out/bench_hits16.log
