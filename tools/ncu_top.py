#!/usr/bin/env python3
"""Summarise an .ncu-rep: key raw metrics + hottest SASS instructions with their dominant stall reasons."""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct", "smsp__issue_active.avg.pct", "sm__inst_executed.avg.per_cycle_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "launch__registers_per_thread ",
        "sm__cycles_elapsed.avg ", "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum "]
for h, u, v in zip(raw[0], raw[1], raw[2]):
    if any((h + " ").startswith(k) or k.strip() in h and k.endswith("pct") for k in keys):
        print("%-80s %-12s %s" % (h, u, v))
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, data = src[1], src[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = {k: sum(f(r, k) for r in data) for k in stalls}
print("stall totals:", sorted(((k, round(100 * v / tot, 1)) for k, v in agg.items()), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:n]:
    s = sorted(((k, f(r, k)) for k in stalls), key=lambda kv: -kv[1])[:2]
    print(r[ix["Address"]][-5:], "%6.2f%%" % (100 * f(r, "# Samples") / tot), "exec", int(f(r, "Instructions Executed")), r[ix["Source"]][:64], s)
