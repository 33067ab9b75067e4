#!/usr/bin/env bash
# Round 2, third GPU call: lean issuer (stage-unrolled, inline spins) against the fast-epilogue-only build and round 1; trace + phase; parity suite
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in ${VARIANTS:-r1 fast lean}; do
  echo "== $v"; B200_BENCH_DIAG=1 B200SCAN_LIB=$PWD/blamm_b200/lib/variants/$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ms/step %.3f  kernel_ms %.3f  value %.3e  e2e %.3e  cand %d hits %d clocks %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value'], d['e2e']['value'], d['config']['candidates_per_step'], d['config']['hits_per_step'], d['clocks']['sm_mhz']))
    elif 'rror' in l: print('   ', l.strip()[:300])
"
done 2>&1 | tee gpurun_out/r2_variants_call3.log
timeout 300 python tools/tc_trace.py 128 10 12 8 > gpurun_out/r2_trace_i8.log 2>&1; tail -7 gpurun_out/r2_trace_i8.log
timeout 300 python tools/tc_phase.py 50 > gpurun_out/r2_phase.log 2>&1; cat gpurun_out/r2_phase.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
