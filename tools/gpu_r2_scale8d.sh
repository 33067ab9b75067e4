#!/usr/bin/env bash
# Round 2: configs[2] through the CLI at -g 8 / 4 / 2 / 1 on one 8-GPU box (same FASTA set), after the CUDA_VISIBLE_DEVICES trimming was removed.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); p = d['path']['phases_s']
        print('%s -g %d: wall %.2f s  e2e %.3e  value %.3e  cuda wait %.2f  format %.2f  write %.2f  kernels/GPU %s' % (sys.argv[1], d['n_gpus'], d['e2e']['wall_s'], d['e2e']['value'], d['value'],
              p.get('setup: wait for the CUDA driver (device count)', 0), p.get('format ordered hits (wall)', 0), p.get('file write (wall)', 0), [round(x) for x in d['path']['per_gpu_kernel_ms']]))
PY
}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29551"
timeout 1200 $TR --nproc-per-node 8 bench.py --gpus 8 --config c3 --steps 2 --warmup 1 --reuse /dev/shm/c3full > gpurun_out/r2_scale8_c3_g8.json 2> gpurun_out/r2_scale8_c3_g8.err; echo "c3 -g 8 rc=$?"; show gpurun_out/r2_scale8_c3_g8.json
timeout 900 python bench.py --gpus 1 --config c3 --steps 2 --warmup 0 --reuse /dev/shm/c3full > gpurun_out/r2_scale8_c3_g1.json 2> gpurun_out/r2_scale8_c3_g1.err; echo "c3 -g 1 rc=$?"; show gpurun_out/r2_scale8_c3_g1.json
timeout 900 $TR --nproc-per-node 2 bench.py --gpus 2 --config c3 --steps 1 --warmup 0 --reuse /dev/shm/c3full > gpurun_out/r2_scale8_c3_g2.json 2> gpurun_out/r2_scale8_c3_g2.err; echo "c3 -g 2 rc=$?"; show gpurun_out/r2_scale8_c3_g2.json
timeout 900 $TR --nproc-per-node 4 bench.py --gpus 4 --config c3 --steps 1 --warmup 0 --reuse /dev/shm/c3full > gpurun_out/r2_scale8_c3_g4.json 2> gpurun_out/r2_scale8_c3_g4.err; echo "c3 -g 4 rc=$?"; show gpurun_out/r2_scale8_c3_g4.json
cat /sys/kernel/mm/transparent_hugepage/shmem_enabled /sys/kernel/mm/transparent_hugepage/enabled 2>/dev/null
rm -rf /dev/shm/c3full
