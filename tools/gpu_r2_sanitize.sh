#!/usr/bin/env bash
# compute-sanitizer (memcheck, then racecheck on the shared-memory kernels) over smoke() and the tests of round 2's new kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r2_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/r2_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -x -q -m gpu -k "ordered_hit_records or hit_buffer_overflow or compact_hit or edge or empty or many_short" > gpurun_out/r2_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -4 gpurun_out/r2_memcheck_tests.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -x -q -m gpu -k "ordered_hit_records" > gpurun_out/r2_racecheck_tests.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_racecheck_tests.log
