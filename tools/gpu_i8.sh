# INT8-operand filter: parity subset, then bench lines per operand kind
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "${K:-random_cases or soft_masked or accumulator_type or golden or edge_blocks or max_length or lower_case}" > gpurun_out/pytest_i8.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_i8.log
for acc in ${ACCS:-0 16}; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --acc $acc > gpurun_out/bench_acc$acc.log 2>&1; echo "bench acc=$acc rc=$?"; tail -1 gpurun_out/bench_acc$acc.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('value %.3e  ms/step %.2f  kernel_ms %.2f  e2e %.3e  cand %d hits %d stages %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['config']['candidates_per_step'], d['config']['hits_per_step'], d['e2e']['stages_ms']))"
done
