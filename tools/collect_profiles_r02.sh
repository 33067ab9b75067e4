#!/usr/bin/env bash
# gpurun_out/r2_* (the GPU calls of round 2) -> profiles/r02_*: what DESIGN.md cites.  usage: tools/collect_profiles_r02.sh
set -e
cd "$(dirname "$0")/.."
G=gpurun_out; P=profiles
cp_if() { [ -f "$1" ] && cp "$1" "$2" || echo "missing $1"; }
tail -n 1 $G/r2_i8_peak.jsonl > $P/r02_i8_peak.json
for v in c2 c2_hits12 c2_bias c2_softmask c4 ref ref_c3; do cp_if $G/r2_bench_$v.json $P/r02_bench_$v.json; done
for f in c2_n8 c2_n8_ascii c2_n4 c2_n1 c3_g8 c3_g1 c5_g8 c4_n8; do cp_if $G/r2_scale8_$f.json $P/r02_scale8_$f.json; done
cp_if $G/r2_scale8_box.txt $P/r02_scale8_box.txt
for f in c2 c3 c3_g1 c4 c5; do cp_if $G/r2_two_$f.json $P/r02_two_$f.json; done
cp_if $G/r2_parity.log $P/r02_parity.log; cp_if $G/r2_parity_exceptions.txt $P/r02_parity_exceptions.txt
cp_if $G/r2_cli_e2e.log $P/r02_cli_e2e.log; cp_if $G/r2_cli_startup.log $P/r02_cli_startup.log
cp_if $G/r2_pytest_gpu.log $P/r02_pytest_gpu.log
cat $G/r2_variants1.log $G/r2_variants_call3.log > $P/r02_variants.log 2>/dev/null || true
# strip torchrun banners from the json files (keep the JSON line only)
for f in $P/r02_bench_*.json $P/r02_scale8_*.json $P/r02_two_*.json; do grep '^{"' "$f" > "$f.tmp" 2>/dev/null && mv "$f.tmp" "$f" || rm -f "$f.tmp"; done
ls -la $P | grep r02
