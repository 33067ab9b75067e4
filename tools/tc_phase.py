#!/usr/bin/env python
"""Where the filter kernel's warps spend their cycles: per-role totals from the B200_PHASE build (make
blamm_b200/lib/libb200scan_phase.so), bench.py's workload, one launch.  usage: python tools/tc_phase.py [Mbp]"""
import ctypes, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200SCAN_LIB", os.path.join(ROOT, "blamm_b200", "lib", "libb200scan_phase.so"))
from blamm_b200 import capi
import bench

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 50.0
n = int(mbp * 1e6)
ms, P, col_len, thr, seq, bg = bench.build_inputs(tempfile.mkdtemp(), n, 0)
sc = capi.Scanner(0, max_block_nt=n + 64, max_hits=max(1 << 20, int(2.2e-4 * n * len(col_len))))
sc.set_engine(capi.ENGINE_TENSOR)
sc.set_motifs(P, col_len, thr)
L = capi.scan_lib(); L.b200scan_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_uint64 * 64)()
hits, t = sc.scan(seq)
assert L.b200scan_debug_trace(sc._ctx, buf, 64) == 0
a = np.frombuffer(buf, dtype=np.uint64).astype(np.float64).reshape(4, 16)
print("score kernel %.2f ms (phase build; the product build has no clock reads), %d hits" % (t["score_ms"], len(hits)))
d = sc.describe()
tiles_per_sm = a[1, 3] / d["sm_count"]
print("column tiles %d, %s; window tiles issued %.0f = %.0f per SM -> %.0f cycles of kernel time per tile at 1.965 GHz" % (
    d["n_tiles"], sc.tensor_info(), a[1, 3], tiles_per_sm, t["score_ms"] * 1e-3 * 1.965e9 / tiles_per_sm))
names = {0: ("producer warp 0 (per 4-tile stage)", ["wait for a free E stage", "fill"]),
         1: ("issuer (per tile)", ["wait for E stages", "wait for the TMEM buffer", "issue MMAs + commits"]),
         2: ("epilogue warp, group 0 (per tile of its buffer = every 2nd tile)", ["wait for tFull", "tcgen05.ld phase until release", "sign compaction + push"])}
for role, (title, phases) in names.items():
    cnt = a[role, 3]
    if cnt == 0:
        print(title, ": no samples"); continue
    tot = sum(a[role, k] for k in range(len(phases)))
    print("%s: %.0f cycles per unit" % (title, tot / cnt))
    for k, ph in enumerate(phases):
        print("    %-36s %7.0f cycles  %5.1f %%" % (ph, a[role, k] / cnt, 100 * a[role, k] / tot))
