#!/usr/bin/env bash
# Round 2, first GPU call: box probe, dense INT8 / FP16 tcgen05 peak (tools/micro/i8_peak.cu), filter-kernel variants on the
# bench workload (tools/variants.sh run), the GPU parity suite on the new default kernel.
#   here:  nvcc ... -o blamm_b200/lib/i8_peak tools/micro/i8_peak.cu ; variants under blamm_b200/lib/variants/*.so
#   box:   gpurun --timeout 1500 -- 'bash tools/gpu_r2_call1.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{ nproc; free -g | head -2; df -h /tmp /dev/shm . | cat; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread"; nvidia-smi -L; } > gpurun_out/r2_box.txt 2>&1
cat gpurun_out/r2_box.txt
for c in 4000 400000; do timeout 120 blamm_b200/lib/i8_peak $c | tee -a gpurun_out/r2_i8_peak.jsonl; done
for v in r1 fast nofast r1; do
  echo "== $v"; B200_BENCH_DIAG=1 B200SCAN_LIB=$PWD/blamm_b200/lib/variants/$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ms/step %.3f  kernel_ms %.3f  value %.3e  e2e %.3e  cand %d hits %d clocks %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['value'], d['e2e']['value'], d['config']['candidates_per_step'], d['config']['hits_per_step'], d['clocks']))
    elif 'rror' in l: print('   ', l.strip()[:300])
"
done 2>&1 | tee gpurun_out/r2_variants1.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
