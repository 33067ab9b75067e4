#!/usr/bin/env bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('%-16s value %.3e ms/step %.2f e2e %.3e (%.2f ms)' % (sys.argv[2], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), {k: round(v, 3) for k, v in d['e2e']['stages_ms_last_block'].items()})
    elif 'rror' in l: print(l[:200])
PY
}
timeout 300 python bench.py --no-cpu-baseline --steps 5 --packed > gpurun_out/tmp_v.json 2>&1; show gpurun_out/tmp_v.json c2
timeout 300 python bench.py --no-cpu-baseline --steps 5 --packed --softmask 0.5 > gpurun_out/tmp_v.json 2>&1; show gpurun_out/tmp_v.json softmask
timeout 600 python -m pytest tests -x -q -m gpu -k "random_cases or soft_masked or golden or overflow or ordered or int8 or mixed or config4 or edge" > gpurun_out/r2_pytest_part.log 2>&1; tail -3 gpurun_out/r2_pytest_part.log
python __graft_entry__.py smoke 2>&1 | tail -4
