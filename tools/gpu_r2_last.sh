#!/usr/bin/env bash
# Round 2, last GPU call: the GPU parity suite and a short bench on the final tree (after the folding functions moved to file
# scope and the host parser / histogram writer were rewritten).  ~4 minutes of box time.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 230 python -m pytest tests -x -q -m gpu > gpurun_out/r2_last_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_last_pytest_gpu.log
timeout 130 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_last_bench.json 2> gpurun_out/r2_last_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2_last_bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print('c2: ms/step %.3f value %.3e e2e %.3e (%.2f ms) frac %.3f pipe %.3f launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['pipe']['frac'], d['gpu_launches']))
PY
tail -n 3 gpurun_out/r2_last_bench.err
