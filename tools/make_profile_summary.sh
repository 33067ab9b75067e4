#!/usr/bin/env bash
# gpurun_out/{launches.csv,prof_filter.ncu-rep,prof_small.ncu-rep} (tools/gpu_prof.sh) -> profiles/rNN_*  usage: tools/make_profile_summary.sh r01
set -e
cd "$(dirname "$0")/.."
R=${1:-r01}
cp gpurun_out/launches.csv profiles/${R}_launches_bench.csv
{ echo "# Round ${R#r} ncu summaries (B200, bench.py default workload: 1800 columns x 100 Mbp, INT8-operand tensor filter: tcgen05.mma.kind::i8, S32 accumulators read back with pack::16b)"
echo "# commands: tools/gpu_prof.sh; raw metrics via 'ncu -i <rep> --page raw --csv', hot SASS via tools/ncu_top.py"
echo; echo "## launch list of 'python bench.py --steps 2 --warmup 3' (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised -> compare SHARES)"
python3 - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[h]; data=rows[h+1:]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in data:
    if len(r)<=vi: continue
    agg.setdefault(r[ki].split('(')[0],[]).append(float(r[vi].replace(',','')))
tot=sum(sum(v) for v in agg.values())
for k,v in agg.items(): print('%-40s launches=%3d  mean %8.3f ms  share of GPU time %5.1f%%'%(k,len(v),sum(v)/len(v)/1e6,100*sum(v)/tot))
PY
echo; echo "## filter_tc_kernel<ACC16 = true, ZMASK = false, PAIR = false, I8 = true> (ncu --set full --clock-control none), one launch = 100 Mbp x 1800 columns"
python3 tools/ncu_top.py gpurun_out/prof_filter.ncu-rep 25
ncu -i gpurun_out/prof_filter.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
for h,u,v in zip(rows[0],rows[1],rows[2]):
    if h in ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__shared_mem_per_block_dynamic','launch__block_size','launch__grid_size','smsp__sass_inst_executed_op_tmem_ldt.sum'): print('%-80s %-12s %s'%(h,u,v))
"
echo; echo "## ordering kernels (bucket_scan, bucket_scatter, bucket_order) and the fused rescorer (rescore_tile_kernel) (ncu --set full, same workload)"
ncu -i gpurun_out/prof_small.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    d=dict(zip(hdr,r)); u=dict(zip(hdr,rows[1]))
    print(d['Kernel Name'].split('(')[0])
    for k in want: print('   %-72s %-10s %s'%(k,u.get(k,''),d.get(k,'')))
"; } > profiles/${R}_ncu_summary.txt
wc -l profiles/${R}_ncu_summary.txt
# DRAM traffic of the dominant kernel -> profiles/${R}_filter_traffic.json (bench.py reports it as roofline.traffic)
ncu -i gpurun_out/prof_filter.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import csv, sys, json
rows = list(csv.reader(sys.stdin)); d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
def b(k): return float(d[k].replace(',', '')) * scale[u[k]]
rd, wr = b('dram__bytes_read.sum'), b('dram__bytes_write.sum')
t = float(d['gpu__time_duration.sum'].replace(',', '')) * {'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}[u['gpu__time_duration.sum']]
json.dump({'kernel': d['Kernel Name'].split('(')[0], 'workload': 'bench.py default (1800 columns x 100 Mbp)', 'dram_bytes_read': rd, 'dram_bytes_write': wr,
           'gpu_time_ms_under_ncu': t, 'tensor_pipe_active_pct': float(d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']),
           'source': 'ncu --set full --clock-control none, profiles/${R}_ncu_summary.txt', 'traffic_bytes_per_launch': rd + wr}, open('profiles/${R}_filter_traffic.json', 'w'), indent=1)
print(open('profiles/${R}_filter_traffic.json').read())
"
