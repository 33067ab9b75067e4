#!/usr/bin/env bash
# BASELINE.json configs[2] in miniature: several manifest groups with their own GC content, N runs inside the records and
# soft-masked stretches; dict + hist + scan by blamm-b200 and by the reference binary, occurrence sets compared.
# usage: tools/cli_groups.sh [groups] [Mbp per group] [motifs]
set -e
cd "$(dirname "$0")/.."
ROOT=$PWD; G=${1:-6}; MBP=${2:-2}; NM=${3:-300}
W=$(mktemp -d); cd $W
python - <<PY
import sys; sys.path.insert(0, "$ROOT")
import numpy as np
from blamm_b200 import synth
synth.make_jaspar_like("motifs.jaspar", $NM, 2024)
rng = np.random.default_rng(31)
n = int($MBP * 1e6)
with open("genomes.mf", "w") as mf:
    for g in range($G):
        gc = 0.36 + 0.12 * g / max(1, $G - 1)
        seq = synth.random_acgt(n, 500 + g, (0.5 - gc / 2, gc / 2, gc / 2, 0.5 - gc / 2))
        for _ in range(40):                                  # N runs and soft-masked stretches
            a = int(rng.integers(0, n - 5000)); seq[a:a + int(rng.integers(1, 3000))] = ord("N")
            b = int(rng.integers(0, n - 5000)); seq[b:b + int(rng.integers(1, 5000))] |= 0x20
        q = n // 3
        synth.write_fasta("g%02d.fa" % g, [("g%02d_chr%d" % (g, i + 1), seq[i * q:(i + 1) * q]) for i in range(3)])
        mf.write("group%02d\tg%02d.fa\n" % (g, g))
PY
B=$ROOT/blamm_b200/lib/blamm-b200; R=$ROOT/oracle/_ref/blamm
export OPENBLAS_NUM_THREADS=1 BLAMM_B200_TIMING=1
t() { local s=$(date +%s.%N); "$@" > log.txt 2>&1 || { cat log.txt; exit 1; }; python -c "print('%.2f' % ($(date +%s.%N) - $s))"; }
mkdir b r; for d in b r; do cp genomes.mf motifs.jaspar g*.fa $d/; done
( cd b; echo "b200      dict $(t $B dict genomes.mf) s, hist $(t $B hist motifs.jaspar genomes.mf) s, scan $(t $B scan -rc -pt 0.0001 motifs.jaspar genomes.mf) s"; grep timing log.txt || true )
( cd r; echo "reference dict $(t $R dict genomes.mf) s, hist $(t $R hist motifs.jaspar genomes.mf) s, scan -t $(nproc) $(t $R scan -rc -pt 0.0001 -t $(nproc) motifs.jaspar genomes.mf) s" )
cmp b/genomes.mf.dict r/genomes.mf.dict && echo "dict files identical"
n=0; bad=0; for f in r/hist_*.dat; do n=$((n+1)); cmp -s $f b/$(basename $f) || bad=$((bad+1)); done; echo "theoretical histogram files: $n, differing: $bad"
cmp b/PWMthresholds.txt r/PWMthresholds.txt && echo "PWMthresholds.txt identical"
python - <<PY
def load(f):
    d = {}
    for l in open(f):
        c = l.rstrip("\n").split("\t")
        d[(c[0], c[2], c[3], c[4], c[6])] = float(c[5])
    return d
a, b = load("b/occurrences.txt"), load("r/occurrences.txt")
common = set(a) & set(b)
print("occurrences: b200 %d reference %d, identical sets: %s (only-b200 %d, only-ref %d)" % (len(a), len(b), set(a) == set(b), len(set(a) - set(b)), len(set(b) - set(a))))
print("max |score diff| / max(1,|s|): %.2g" % max((abs(a[k] - b[k]) / max(1.0, abs(b[k])) for k in common), default=0.0))
PY
cd /; rm -rf $W
