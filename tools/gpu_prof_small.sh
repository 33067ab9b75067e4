# refresh of the launch list and of the small kernels' captures (the filter kernel's capture stays: tools/gpu_prof.sh)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'expand_kernel|rescore_kernel|pack_ascii' -s 6 -c 3 -f -o gpurun_out/prof_small python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_small.log 2>&1
ls -la gpurun_out
