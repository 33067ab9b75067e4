#!/usr/bin/env bash
# Round 2, call 9: bench with the hand-over calibration (default, forced both ways, background-biased sequence), c5 at a tenth.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('%s N=%d: value %.3e e2e %.3e ms/step %.2f e2e ms %.2f' % (sys.argv[1], d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e'].get('ms_per_step') or 0),
              d['path'].get('hand_over', '')[:40], d['path'].get('hand_over_calibration'), d['e2e'].get('stages_ms_last_block'))
PY
}
timeout 600 python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "rc=$?"; show gpurun_out/r2_bench_c2.json; tail -n 3 gpurun_out/r2_bench_c2.err
timeout 300 python bench.py --no-cpu-baseline --ascii --steps 5 > gpurun_out/r2_bench_c2_ascii.json 2>&1; show gpurun_out/r2_bench_c2_ascii.json
timeout 300 python bench.py --no-cpu-baseline --packed --steps 5 > gpurun_out/r2_bench_c2_packed.json 2>&1; show gpurun_out/r2_bench_c2_packed.json
timeout 300 python bench.py --no-cpu-baseline --bias > gpurun_out/r2_bench_c2_bias.json 2>&1; show gpurun_out/r2_bench_c2_bias.json
timeout 300 python bench.py --no-cpu-baseline --softmask 0.5 --steps 5 > gpurun_out/r2_bench_c2_softmask.json 2>&1; show gpurun_out/r2_bench_c2_softmask.json
timeout 900 python bench.py --config c5 --gbp 0.31 --steps 1 --warmup 0 > gpurun_out/r2_bench_c5_small.json 2> gpurun_out/r2_bench_c5_small.err; echo "c5 rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_c5_small.json'):
    if l.startswith('{'):
        d = json.loads(l); print('c5 small: hist -e %.2f s, scan wall %.2f s' % (d['path']['hist_s'], d['e2e']['wall_s']))
PY
