#!/usr/bin/env bash
# Round 2, second GPU call: per-tile timeline (B200_TRACE build) and per-role cycle totals (B200_PHASE build) of the new epilogue
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/tc_trace.py 128 10 12 8 > gpurun_out/r2_trace_i8.log 2>&1; tail -32 gpurun_out/r2_trace_i8.log
timeout 300 python tools/tc_phase.py 50 > gpurun_out/r2_phase.log 2>&1; cat gpurun_out/r2_phase.log
