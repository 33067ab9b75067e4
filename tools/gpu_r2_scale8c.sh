#!/usr/bin/env bash
# Round 2, final 8-GPU call: c2 at N = 8 / 1 and c3 (-g 8, -g 1) + c5 with the final kernels and writer.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('%s N=%d: value %.3e e2e %.3e ms/step %.2f e2e %s' % (sys.argv[1], d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e'].get('ms_per_step') or d['e2e'].get('wall_s')),
              d['path'].get('hand_over', '')[:30], d['path'].get('hand_over_calibration'), d['path'].get('hist_s'))
PY
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_scale8_c2_n8.json 2> gpurun_out/r2_scale8_c2_n8.err; echo "c2 N=8 rc=$?"; show gpurun_out/r2_scale8_c2_n8.json
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale8_c2_n1.json 2> gpurun_out/r2_scale8_c2_n1.err; echo "c2 N=1 rc=$?"; show gpurun_out/r2_scale8_c2_n1.json
timeout 1200 $TR --nproc-per-node 8 bench.py --gpus 8 --config c3 --steps 2 --warmup 1 --reuse /dev/shm/c3full > gpurun_out/r2_scale8_c3_g8.json 2> gpurun_out/r2_scale8_c3_g8.err; echo "c3 -g 8 rc=$?"; show gpurun_out/r2_scale8_c3_g8.json; tail -n 2 gpurun_out/r2_scale8_c3_g8.err
timeout 900 python bench.py --gpus 1 --config c3 --steps 1 --warmup 1 --reuse /dev/shm/c3full > gpurun_out/r2_scale8_c3_g1.json 2> gpurun_out/r2_scale8_c3_g1.err; echo "c3 -g 1 rc=$?"; show gpurun_out/r2_scale8_c3_g1.json; tail -n 2 gpurun_out/r2_scale8_c3_g1.err
timeout 900 $TR --nproc-per-node 8 bench.py --gpus 8 --config c5 --steps 1 --warmup 0 --reuse /dev/shm/c3full > gpurun_out/r2_scale8_c5_g8.json 2> gpurun_out/r2_scale8_c5_g8.err; echo "c5 -g 8 rc=$?"; show gpurun_out/r2_scale8_c5_g8.json; tail -n 2 gpurun_out/r2_scale8_c5_g8.err
rm -rf /dev/shm/c3full
