#!/usr/bin/env bash
# 16-byte raw entries + fused rescorer only: parity suite, bench (uniform, soft-masked, c4), launch list, full ncu of filter + small kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_gpu.log
show() { python - "$1" "$2" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('%-16s value %.3e ms/step %.2f e2e %.3e (%.2f ms)' % (sys.argv[2], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), {k: round(v, 3) for k, v in d['e2e']['stages_ms_last_block'].items()})
    elif 'rror' in l: print(l[:200])
PY
}
timeout 600 python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; show gpurun_out/r2_bench_c2.json c2; tail -n 2 gpurun_out/r2_bench_c2.err
timeout 300 python bench.py --no-cpu-baseline --steps 5 --softmask 0.5 > gpurun_out/r2_bench_c2_softmask.json 2>&1; show gpurun_out/r2_bench_c2_softmask.json softmask
timeout 300 python bench.py --no-cpu-baseline --bias > gpurun_out/r2_bench_c2_bias.json 2>&1; show gpurun_out/r2_bench_c2_bias.json bias
timeout 300 python bench.py --no-cpu-baseline --hits 12 --steps 5 > gpurun_out/r2_bench_c2_hits12.json 2>&1; show gpurun_out/r2_bench_c2_hits12.json hits12
timeout 600 python bench.py --config c4 --mbp 320 --steps 2 > gpurun_out/r2_bench_c4.json 2>&1; show gpurun_out/r2_bench_c4.json c4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --packed > gpurun_out/r2_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_tc -s 6 -c 1 -f -o gpurun_out/prof_filter python bench.py --steps 1 --warmup 3 --no-cpu-baseline --packed > gpurun_out/r2_ncu_filter.log 2>&1; echo "ncu filter rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bucket_|rescore_tile' -s 15 -c 5 -f -o gpurun_out/prof_small python bench.py --steps 1 --warmup 3 --no-cpu-baseline --packed > gpurun_out/r2_ncu_small.log 2>&1; echo "ncu small rc=$?"
