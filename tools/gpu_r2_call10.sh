#!/usr/bin/env bash
# Round 2, call 10: fused expand + rescore (rescore_tile_kernel) -- parity suite, bench against the round-1 chain (B200SCAN_RESCORE=list), launch list.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest_gpu.log
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print('%s N=%d: value %.3e e2e %.3e ms/step %.2f e2e ms %.2f' % (sys.argv[1], d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e'].get('ms_per_step') or 0), d['e2e'].get('stages_ms_last_block'), d['path']['candidates_last_block'])
PY
}
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_c2_fused.json 2> gpurun_out/r2_bench_c2_fused.err; echo "rc=$?"; show gpurun_out/r2_bench_c2_fused.json; tail -n 3 gpurun_out/r2_bench_c2_fused.err
B200SCAN_RESCORE=list timeout 600 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2_bench_c2_list.json 2>&1; show gpurun_out/r2_bench_c2_list.json
timeout 600 python bench.py --no-cpu-baseline --softmask 0.5 --steps 5 > gpurun_out/r2_bench_c2_softmask.json 2>&1; show gpurun_out/r2_bench_c2_softmask.json
timeout 600 python bench.py --no-cpu-baseline --config c4 --mbp 320 --steps 2 > gpurun_out/r2_bench_c4_small.json 2>&1; show gpurun_out/r2_bench_c4_small.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --packed > gpurun_out/r2_bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r2_launches.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[h]; data=rows[h+1:]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in data:
    if len(r)<=vi: continue
    agg.setdefault(r[ki].split('(')[0],[]).append(float(r[vi].replace(',','')))
for k,v in agg.items(): print('%-60s launches=%3d  mean %8.3f ms'%(k[:60],len(v),sum(v)/len(v)/1e6))
PY
