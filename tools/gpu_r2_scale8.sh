#!/usr/bin/env bash
# Round 2, the 8-GPU call (gpurun --gpus 8): BASELINE.json's configs at full size.
#   c2 at N = 8 and N = 1 on the same box (what the driver's SCALE run measures),
#   c3: 3.1 Gbp x 24 groups through the CLI with -g 8 and -g 1 on the same FASTA set (FASTA in -> occurrences.txt out),
#   c5: `hist -e` over every whole group on 8 GPUs, then scan,   c4: 10,000 PWMs x 1 Gbp dealt to 8 ranks.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi -L; nvidia-smi topo -m; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"; df -h /dev/shm /tmp; } > gpurun_out/r2_scale8_box.txt 2>&1
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('%s N=%d: value %.3e e2e %.3e ms/step %.2f' % (sys.argv[1], d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step']),
              {k: d['e2e'].get(k) for k in ('ms_per_step', 'wall_s', 'host_pack_ms_per_step')}, d.get('clocks', {}).get('reasons'))
PY
}
timeout 600 $TR8 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_scale8_c2_n8.json 2> gpurun_out/r2_scale8_c2_n8.err; echo "c2 N=8 rc=$?"; show gpurun_out/r2_scale8_c2_n8.json
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale8_c2_n1.json 2> gpurun_out/r2_scale8_c2_n1.err; echo "c2 N=1 rc=$?"; show gpurun_out/r2_scale8_c2_n1.json
timeout 1200 $TR8 bench.py --gpus 8 --config c3 --steps 2 --warmup 1 --reuse /dev/shm/c3full > gpurun_out/r2_scale8_c3_g8.json 2> gpurun_out/r2_scale8_c3_g8.err; echo "c3 -g 8 rc=$?"; show gpurun_out/r2_scale8_c3_g8.json; tail -n 3 gpurun_out/r2_scale8_c3_g8.err
timeout 900 python bench.py --gpus 1 --config c3 --steps 1 --warmup 0 --reuse /dev/shm/c3full > gpurun_out/r2_scale8_c3_g1.json 2> gpurun_out/r2_scale8_c3_g1.err; echo "c3 -g 1 rc=$?"; show gpurun_out/r2_scale8_c3_g1.json; tail -n 3 gpurun_out/r2_scale8_c3_g1.err
timeout 900 $TR8 bench.py --gpus 8 --config c5 --steps 1 --warmup 0 --reuse /dev/shm/c3full > gpurun_out/r2_scale8_c5_g8.json 2> gpurun_out/r2_scale8_c5_g8.err; echo "c5 -g 8 rc=$?"; show gpurun_out/r2_scale8_c5_g8.json; tail -n 3 gpurun_out/r2_scale8_c5_g8.err
rm -rf /dev/shm/c3full
timeout 900 $TR8 bench.py --gpus 8 --config c4 --steps 3 --warmup 3 > gpurun_out/r2_scale8_c4_n8.json 2> gpurun_out/r2_scale8_c4_n8.err; echo "c4 N=8 rc=$?"; show gpurun_out/r2_scale8_c4_n8.json; tail -n 3 gpurun_out/r2_scale8_c4_n8.err
