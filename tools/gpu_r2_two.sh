#!/usr/bin/env bash
# Round 2: two-GPU validation of every multi-GPU path (gpurun --gpus 2): torchrun bench (c2, c4), the CLI with -g 2 (c3, c5 at a tenth),
# and the new > 2^31-offset test.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 600 python -m pytest tests -x -q -m gpu -k "beyond_2_to_31 or ordered_hit or config3" > gpurun_out/r2_two_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_two_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_two_c2.json 2> gpurun_out/r2_two_c2.err; echo "c2 N=2 rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2_two_c2.json'):
    if l.startswith('{'):
        d = json.loads(l); print('c2 N=%d: ms/step %.3f value %.3e e2e %.3e (%.2f ms/step, pack %.2f ms)' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['host_pack_ms_per_step']), d['path']['hits_per_step'])
PY
tail -3 gpurun_out/r2_two_c2.err
timeout 900 $TR bench.py --gpus 2 --config c3 --gbp 0.31 --steps 2 --warmup 1 --reuse /dev/shm/c3small > gpurun_out/r2_two_c3.json 2> gpurun_out/r2_two_c3.err; echo "c3 -g 2 rc=$?"
timeout 900 python bench.py --gpus 1 --config c3 --gbp 0.31 --steps 2 --warmup 1 --reuse /dev/shm/c3small > gpurun_out/r2_two_c3_g1.json 2> gpurun_out/r2_two_c3_g1.err; echo "c3 -g 1 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2_two_c3.json', 'gpurun_out/r2_two_c3_g1.json'):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); print('c3 small -g %d: value %.3e e2e %.3e wall %.2f s matches %d' % (d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['wall_s'], d['path']['matches']), d['path']['per_gpu_kernel_ms'])
PY
tail -3 gpurun_out/r2_two_c3.err gpurun_out/r2_two_c3_g1.err
timeout 900 $TR bench.py --gpus 2 --config c5 --gbp 0.31 --steps 1 --warmup 0 --reuse /dev/shm/c3small > gpurun_out/r2_two_c5.json 2> gpurun_out/r2_two_c5.err; echo "c5 -g 2 rc=$?"; tail -c 600 gpurun_out/r2_two_c5.json; tail -3 gpurun_out/r2_two_c5.err
timeout 900 $TR bench.py --gpus 2 --config c4 --mbp 320 --steps 2 --warmup 3 > gpurun_out/r2_two_c4.json 2> gpurun_out/r2_two_c4.err; echo "c4 N=2 rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2_two_c4.json'):
    if l.startswith('{'):
        d = json.loads(l); print('c4 N=%d: ms/step %.3f value %.3e e2e %.3e frac %.3f' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac']))
PY
tail -3 gpurun_out/r2_two_c4.err
