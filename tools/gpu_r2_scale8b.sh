#!/usr/bin/env bash
# Round 2, second 8-GPU call: c2 at N = 8 / 4 with the calibrated hand-over (and both forced), c5 with the finer hist -e chunks.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('%s N=%d: value %.3e e2e %.3e ms/step %.2f e2e ms %s' % (sys.argv[1], d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e'].get('ms_per_step') or d['e2e'].get('wall_s')),
              d['path'].get('hand_over', '')[:30], d['path'].get('hand_over_calibration'), d['path'].get('hist_s'))
PY
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_scale8_c2_n8.json 2> gpurun_out/r2_scale8_c2_n8.err; echo "c2 N=8 rc=$?"; show gpurun_out/r2_scale8_c2_n8.json
timeout 600 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 10 --warmup 3 --ascii > gpurun_out/r2_scale8_c2_n8_ascii.json 2> gpurun_out/r2_scale8_c2_n8_ascii.err; echo "c2 N=8 ascii rc=$?"; show gpurun_out/r2_scale8_c2_n8_ascii.json
timeout 600 $TR --nproc-per-node 4 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_scale8_c2_n4.json 2> gpurun_out/r2_scale8_c2_n4.err; echo "c2 N=4 rc=$?"; show gpurun_out/r2_scale8_c2_n4.json
timeout 900 $TR --nproc-per-node 8 bench.py --gpus 8 --config c5 --steps 1 --warmup 0 > gpurun_out/r2_scale8_c5_g8.json 2> gpurun_out/r2_scale8_c5_g8.err; echo "c5 -g 8 rc=$?"; show gpurun_out/r2_scale8_c5_g8.json; tail -n 3 gpurun_out/r2_scale8_c5_g8.err
