#!/usr/bin/env bash
# Race detection (default) or memory / undefined-behaviour checks (SAN=address) for the CLI's host pipeline, no GPU: blamm-b200
# built with -fsanitize=thread (SAN=address: -fsanitize=address,undefined) and linked against the CPU suite's
# test-only stand-in library, driven through the multi-device paths -- scan on 4 devices with small chunks and out-of-order
# completion, the 12-byte records + host sort, chunks refused and scored in halves, the error path, hist -e on 3 devices, the
# theoretical hist, dict on 4 parser threads, the stream-order self-test.  Prints the number of sanitizer reports per run.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SAN=${SAN:-thread}
if [ "$SAN" = address ]; then SANFLAGS="-fsanitize=address,undefined"; else SANFLAGS="-fsanitize=thread"; fi
W=$(mktemp -d /tmp/tsan_cli.XXXXXX)
mkdir -p $ROOT/tests/mock/_build
g++ -O2 -std=c++17 -shared -fPIC $ROOT/tests/mock/mock_b200scan.cpp -o $ROOT/tests/mock/_build/libb200scan.so -L$ROOT/oracle -loracle -Wl,-rpath,$ROOT/oracle -lpthread
g++ -O1 -g -std=c++17 $SANFLAGS -fPIC $ROOT/blamm_b200/host/motifs.cpp $ROOT/blamm_b200/host/sequence.cpp $ROOT/blamm_b200/host/cli.cpp \
    -o $W/blamm-b200.tsan -L$ROOT/tests/mock/_build -lb200scan -Wl,-rpath,$ROOT/tests/mock/_build -lpthread
python - <<PY
import sys
sys.path.insert(0, "$ROOT"); sys.path.insert(0, "$ROOT/tests")
import test_cli_mock as T
T._make_inputs("$W", 204, n_groups=3)
PY
cd $W
B=$W/blamm-b200.tsan
export TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" ASAN_OPTIONS="detect_leaks=0:halt_on_error=0" UBSAN_OPTIONS="halt_on_error=0"
run() { local name=$1; shift; local out; out=$("$@" 2>&1 || true); printf "%-34s %s sanitizer reports (%s); %s\n" "$name" "$(grep -c -E 'WARNING: ThreadSanitizer|ERROR: AddressSanitizer|runtime error:' <<<"$out" || true)" "$SAN" "$(grep -E 'Wrote|bye|identical|error' <<<"$out" | tail -1)"; }
run "dict, 4 parser threads"        env BLAMM_B200_INGEST_THREADS=4 $B dict seq.mf
run "hist (theoretical)"            $B hist -t 4 motifs.jaspar seq.mf
run "scan, 4 devices"               env MOCK_B200SCAN_DEVICES=4 MOCK_B200SCAN_DELAY_US=2000 BLAMM_B200_CHUNK=15000 $B scan -rc -pt 0.001 -t 4 motifs.jaspar seq.mf
run "scan, 12-byte records"         env MOCK_B200SCAN_DEVICES=3 BLAMM_B200_HITS=12 BLAMM_B200_CHUNK=15000 $B scan -rc -pt 0.001 -t 4 motifs.jaspar seq.mf
run "scan, refused chunks halved"   env MOCK_B200SCAN_DEVICES=3 MOCK_B200SCAN_DELAY_US=1000 BLAMM_B200_CHUNK=30000 MOCK_B200SCAN_HIT_BUDGET=2000 $B scan -rc -pt 0.002 -t 4 motifs.jaspar seq.mf
run "scan, error path"              env MOCK_B200SCAN_DEVICES=3 BLAMM_B200_CHUNK=30000 MOCK_B200SCAN_HIT_BUDGET=0 $B scan -rc -pt 0.002 -t 4 motifs.jaspar seq.mf
run "scan, injected collect failure" env MOCK_B200SCAN_DEVICES=3 MOCK_B200SCAN_DELAY_US=500 BLAMM_B200_CHUNK=9000 MOCK_B200SCAN_FAIL_COLLECT=7 $B scan -rc -pt 0.001 -t 4 motifs.jaspar seq.mf
run "scan, injected submit failure"  env MOCK_B200SCAN_DEVICES=3 MOCK_B200SCAN_DELAY_US=500 BLAMM_B200_CHUNK=9000 MOCK_B200SCAN_FAIL_SUBMIT=9 $B scan -rc -pt 0.001 -t 4 motifs.jaspar seq.mf
mkdir -p he
run "hist -e, 3 devices"            env MOCK_B200SCAN_DEVICES=3 BLAMM_B200_CHUNK=9000 $B hist -e -H he -t 4 motifs.jaspar seq.mf
run "selftest-order, 5 workers"     $B selftest-order 24 5 3
run "selftest-writer"               $B selftest-writer 200000 4
rm -rf $W
