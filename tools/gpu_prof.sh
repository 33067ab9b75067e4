set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
MBP=${MBP:-20}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --mbp $MBP --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filter_tc -s 2 -c 1 -f -o gpurun_out/prof_filter python bench.py --mbp $MBP --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
