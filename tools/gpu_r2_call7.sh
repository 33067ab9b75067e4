#!/usr/bin/env bash
# Round 2, call 7: ncu --set full of the small kernels (ordering, rescoring) and of the filter; CLI with the mapped writer; c3 at a tenth.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bucket_|expand_kernel|rescore_kernel' -s 18 -c 6 -f -o gpurun_out/r2_prof_small python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_small.log 2>&1; echo "ncu small rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_tc -s 6 -c 1 -f -o gpurun_out/r2_prof_filter python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_filter.log 2>&1; echo "ncu filter rc=$?"
BLAMM_B200_TIMING=1 timeout 600 bash tools/cli_e2e.sh 100 8 1 > gpurun_out/r2_cli_e2e.log 2>&1; grep -E "scan  |timing|identical|lines" gpurun_out/r2_cli_e2e.log
timeout 900 python bench.py --config c3 --gbp 0.31 --steps 2 --warmup 1 > gpurun_out/r2_bench_c3_small.json 2> gpurun_out/r2_bench_c3_small.err; echo "c3 rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_c3_small.json'):
    if l.startswith('{'):
        d = json.loads(l); print('c3 small: value %.3e e2e %.3e wall %.2f s' % (d['value'], d['e2e']['value'], d['e2e']['wall_s'])); print(json.dumps(d['path']['phases_s'], indent=0))
PY
timeout 400 python bench.py --impl reference --config c3 --warmup 0 --steps 1 > gpurun_out/r2_bench_ref_c3.json 2> gpurun_out/r2_bench_ref_c3.err; echo "ref c3 rc=$?"; cat gpurun_out/r2_bench_ref_c3.json
