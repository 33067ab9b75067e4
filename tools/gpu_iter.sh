# iteration loop on the GPU box: parity tests, bench line, launch list + one full ncu capture of the filter kernel
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
if [ "${DEBUG:-0}" = "1" ]; then timeout 600 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1; echo "tc_debug rc=$?"; grep -E "info|hits|IDENT|MISM|missing|scaled|FAILED" gpurun_out/tc_debug.log | tail -40; fi
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench.log
if [ "${NCU:-1}" = "1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_tc -s 2 -c 1 -f -o gpurun_out/prof_filter python bench.py --mbp 20 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
