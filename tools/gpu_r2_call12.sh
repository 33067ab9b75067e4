#!/usr/bin/env bash
# fused rescorer variants: shared-memory budget (CTAs per SM) against the round-1 chain
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('%-16s value %.3e ms/step %.2f' % (sys.argv[2], d['value'], d['ms_per_step']), {k: round(v, 3) for k, v in d['e2e']['stages_ms_last_block'].items()})
    elif 'rror' in l: print(l[:200])
PY
}
for v in 9728 6400 4864 3584 2560; do B200SCAN_FUSE_MAXW=$v timeout 300 python bench.py --no-cpu-baseline --steps 5 --packed > gpurun_out/tmp_v.json 2>&1; show gpurun_out/tmp_v.json maxw=$v; done
B200SCAN_RESCORE=list timeout 300 python bench.py --no-cpu-baseline --steps 5 --packed > gpurun_out/tmp_v.json 2>&1; show gpurun_out/tmp_v.json list
