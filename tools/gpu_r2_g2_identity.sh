#!/usr/bin/env bash
# The occurrence file and the empirical histograms must not depend on -g: 6 groups x 20 Mbp, scan and hist -e with -g 2 and -g 1, files compared byte for byte.
# gpurun --gpus 2 -- 'bash tools/gpu_r2_g2_identity.sh'
set -e
cd $GRAFT_REPO_ROOT
ROOT=$PWD
W=$(mktemp -d -p /dev/shm); cd $W
python - <<PY
import sys; sys.path.insert(0, "$ROOT")
import numpy as np
from blamm_b200 import synth
synth.make_jaspar_like("motifs.jaspar", 300, 2024)
rng = np.random.default_rng(31)
n = 20_000_000
with open("genomes.mf", "w") as mf:
    for g in range(6):
        gc = 0.36 + 0.12 * g / 5
        seq = synth.random_acgt(n, 500 + g, (0.5 - gc / 2, gc / 2, gc / 2, 0.5 - gc / 2))
        for _ in range(40):
            a = int(rng.integers(0, n - 5000)); seq[a:a + int(rng.integers(1, 3000))] = ord("N")
            b = int(rng.integers(0, n - 5000)); seq[b:b + int(rng.integers(1, 5000))] |= 0x20
        q = n // 3
        synth.write_fasta("g%02d.fa" % g, [("g%02d_chr%d" % (g, i + 1), seq[i * q:(i + 1) * q]) for i in range(3)])
        mf.write("group%02d\tg%02d.fa\n" % (g, g))
PY
B=$ROOT/blamm_b200/lib/blamm-b200
export BLAMM_B200_CHUNK=3000000
$B dict genomes.mf > /dev/null
mkdir h1 h2
$B hist -e -l 30000000 -g 1 -H h1 motifs.jaspar genomes.mf > /dev/null
$B hist -e -l 30000000 -g 2 -H h2 motifs.jaspar genomes.mf > /dev/null
n=0; bad=0; for f in h1/*.dat; do n=$((n+1)); cmp -s $f h2/$(basename $f) || bad=$((bad+1)); done; echo "hist -e: $n files, differing between -g 1 and -g 2: $bad"
$B scan -rc -pt 0.0001 -g 1 -H h1 -o occ_g1.txt motifs.jaspar genomes.mf > /dev/null
$B scan -rc -pt 0.0001 -g 2 -H h1 -o occ_g2.txt --stats stats.json motifs.jaspar genomes.mf > /dev/null
python -c "import json; d=json.load(open('stats.json')); print('chunks per device at -g 2:', [x['chunks'] for x in d['devices']])"
echo "occurrences: $(wc -l < occ_g1.txt) lines; -g 1 vs -g 2: $(cmp -s occ_g1.txt occ_g2.txt && echo byte-identical || echo DIFFERENT)"
BLAMM_B200_HITS=12 $B scan -rc -pt 0.0001 -g 2 -H h1 -o occ_g2_12.txt motifs.jaspar genomes.mf > /dev/null
echo "12-byte hand-over + host sort vs ordered records: $(cmp -s occ_g1.txt occ_g2_12.txt && echo byte-identical || echo DIFFERENT)"
cd /; rm -rf $W
