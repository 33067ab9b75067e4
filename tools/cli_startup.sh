#!/usr/bin/env bash
# Where does a short `blamm-b200 scan` spend its time?  2 Mbp x 1800 columns, phases of the CLI and of b200scan_create.
set -e
cd "$(dirname "$0")/.."
ROOT=$PWD
W=$(mktemp -d); cd $W
python - <<PY
import sys; sys.path.insert(0, "$ROOT")
from blamm_b200 import synth
synth.make_jaspar_like("motifs.jaspar", 900, 2024)
seq = synth.random_acgt(2_000_000, 4242)
synth.write_fasta("g.fa", [("chr1", seq)])
open("g.mf", "w").write("syn\tg.fa\n")
PY
B=$ROOT/blamm_b200/lib/blamm-b200
nvidia-smi --query-gpu=persistence_mode --format=csv,noheader | head -1
$B dict g.mf > /dev/null; $B hist motifs.jaspar g.mf > /dev/null
export BLAMM_B200_TIMING=1 B200SCAN_TIMING=1
for rep in 1 2 3; do
  s=$(date +%s.%N); $B scan -rc -pt 0.0001 motifs.jaspar g.mf > log.txt 2>&1; e=$(date +%s.%N)
  echo "run $rep: wall $(python -c "print('%.2f' % ($e - $s))") s"; grep -E "timing|b200scan_create" log.txt
done
echo "--- python: cuInit + context only"
python - <<PY
import ctypes, time
t0 = time.time(); L = ctypes.CDLL("libcuda.so.1"); L.cuInit(0); t1 = time.time()
d = ctypes.c_int(); L.cuDeviceGet(ctypes.byref(d), 0); c = ctypes.c_void_p(); L.cuDevicePrimaryCtxRetain(ctypes.byref(c), d); t2 = time.time()
print("cuInit %.2f s, primary context %.2f s" % (t1 - t0, t2 - t1))
PY
cd /; rm -rf $W
