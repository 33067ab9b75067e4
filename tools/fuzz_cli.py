#!/usr/bin/env python
"""Seeded fuzz of the blamm-b200 command line against the UNMODIFIED reference binary (oracle/_ref/blamm), without a GPU: the
CLI runs on the CPU suite's test-only stand-in library (tests/mock/mock_b200scan.cpp; build it with `pytest tests/test_cli_mock.py`
or tools/host_pipeline_bench.sh).  Inputs: odd FASTA files (headers with blanks / tabs / long names, empty lines, lines of 0 to 3000
characters, lower case, N and other IUPAC letters, LF or CRLF, with or without a final line end), random motif sets, every
threshold mode; the CLI side with random device counts, chunk sizes, hand-overs and record formats.

  dict : <manifest>.dict byte-identical (and the same exit code)
  scan : the same occurrence text, or north_star's rule through tools/parity_list.py where the reference's sgemm sums in another
         order (identical sets, scores within 1e-4, exceptions only within 1e-4 of their threshold)
  hist : `hist -e` (with -l cuts and -b): every .dat identical, or at most a few observations one bin over for the same reason

usage: fuzz_cli.py dict|scan|hist [runs=100] [first seed=0] [--long]      (--long: 20-60 motifs of up to 35 positions)
"""
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blamm_b200 import synth  # noqa: E402

CLI = os.path.join(ROOT, "blamm_b200", "lib", "blamm-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "blamm")
MOCK = os.path.join(ROOT, "tests", "mock", "_build")


def envs():
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    ref = dict(env)
    ob = os.path.join(ROOT, "oracle", "_ref", "openblas_dir.txt")
    if os.path.exists(ob):
        ref["LD_LIBRARY_PATH"] = open(ob).read().strip() + ":" + env.get("LD_LIBRARY_PATH", "")
    return env, ref, dict(env, LD_LIBRARY_PATH=MOCK + ":" + env.get("LD_LIBRARY_PATH", ""))


def odd_fasta(rng):
    out = []
    for r in range(rng.randint(1, 4)):
        out.append(">" + rng.choice(["s", "chr", "x y", "tab\tsep", "  lead", "a" * rng.randint(1, 30)]) + str(r) + rng.choice(["", " desc", "\tmore words"]))
        for _ in range(rng.randint(0, 40)):
            n = rng.choice([0, 1, 5, 60, 61, 200, rng.randint(1, 3000)])
            alpha = rng.choice(["ACGT", "ACGT", "ACGTacgt", "ACGTN", "ACGTacgtNnRYKM", "AC GT", "ACGT-*"])
            out.append("".join(rng.choice(alpha) for _ in range(n)))
    eol = rng.choice(["\n", "\n", "\r\n"])
    return eol.join(out) + (eol if rng.random() < 0.8 else "")


def inputs(w, rng, seed, long_motifs):
    if long_motifs:
        synth.make_jaspar_like(os.path.join(w, "motifs.jaspar"), rng.randint(20, 60), seed, uniform_len=(5, 35))
    else:
        synth.make_jaspar_like(os.path.join(w, "motifs.jaspar"), rng.randint(1, 12), seed, uniform_len=(5, 14))
    manifest = []
    for f in range(rng.randint(1, 3)):
        open(os.path.join(w, "f%d.fa" % f), "w", newline="").write(odd_fasta(rng))
        manifest.append("g%d\tf%d.fa\n" % (f % 2, f))
    open(os.path.join(w, "seq.mf"), "w").write("".join(manifest))


def run(cmd, cwd, env):
    return subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True)


def main():
    what = sys.argv[1]
    runs = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 100
    first = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("-") else 0
    long_motifs = "--long" in sys.argv
    env, refenv, mockenv = envs()
    base = tempfile.mkdtemp(prefix="fuzz_cli_")
    identical = tolerated = skipped = bad = 0
    for it in range(first, first + runs):
        rng = random.Random(100003 * {"dict": 1, "scan": 2, "hist": 3}[what] + it)
        w = os.path.join(base, "w")
        shutil.rmtree(w, ignore_errors=True)
        os.makedirs(os.path.join(w, "b"))
        inputs(w, rng, it, long_motifs)
        if what == "dict":
            for f in os.listdir(w):
                if os.path.isfile(os.path.join(w, f)):
                    shutil.copy(os.path.join(w, f), os.path.join(w, "b", f))
            r1 = run([REF, "dict", "seq.mf"], w, refenv)
            r2 = run([CLI, "dict", "seq.mf"], os.path.join(w, "b"), dict(env, BLAMM_B200_INGEST_THREADS=str(rng.choice([1, 2, 5]))))
            a = open(os.path.join(w, "seq.mf.dict"), "rb").read() if os.path.exists(os.path.join(w, "seq.mf.dict")) else None
            b = open(os.path.join(w, "b", "seq.mf.dict"), "rb").read() if os.path.exists(os.path.join(w, "b", "seq.mf.dict")) else None
            if r1.returncode == r2.returncode and a == b:
                identical += 1
            else:
                bad += 1
                print("MISMATCH seed", it, r1.returncode, r2.returncode, r2.stderr[-200:])
            continue
        e = dict(mockenv, MOCK_B200SCAN_DEVICES=str(rng.randint(1, 4)), BLAMM_B200_CHUNK=str(rng.choice([1024, 3000, 50000, 1 << 25])),
                 MOCK_B200SCAN_DELAY_US=str(rng.choice([0, 500])))
        if what == "scan":
            mode = rng.choice([["-pt", "0.01"], ["-rt", "0.8"], ["-at", "3"], ["-rc", "-pt", "0.005"], ["-rc", "-at", "4.5"], ["-s", "-rc", "-at", "5"]])
            if rng.random() < 0.3:
                e["BLAMM_B200_HITS"] = "12"
            if rng.random() < 0.3:
                e["BLAMM_B200_ASCII"] = "1"
            ok = all(run([REF] + a, w, refenv).returncode == 0 for a in
                     (["dict", "seq.mf"], ["hist", "motifs.jaspar", "seq.mf"], ["scan", "-t", "2", "-o", "ref.txt"] + mode + ["motifs.jaspar", "seq.mf"]))
            if not ok:
                skipped += 1
                continue
            r = run([CLI, "scan", "-o", "got.txt"] + mode + ["motifs.jaspar", "seq.mf"], w, e)
            if r.returncode != 0:
                bad += 1
                print("FAIL seed", it, mode, r.stderr[-300:])
                continue
            if sorted(open(os.path.join(w, "got.txt")).readlines()) == sorted(open(os.path.join(w, "ref.txt")).readlines()):
                identical += 1
                continue
            m = [x for x in mode if x != "-s"]
            cmd = [sys.executable, os.path.join(ROOT, "tools", "parity_list.py"), "--ours", os.path.join(w, "got.txt"), "--ref", os.path.join(w, "ref.txt"),
                   "--motifs", os.path.join(w, "motifs.jaspar"), "--manifest", os.path.join(w, "seq.mf"), "--histdir", w]
            if "-rc" in m:
                cmd.append("--rc")
                m.remove("-rc")
            q = run(cmd + ["--" + m[0][1:], m[1]], w, env)
            if q.returncode == 0 and "PARITY OK" in q.stdout:
                tolerated += 1
            else:
                bad += 1
                print("MISMATCH seed", it, mode, q.stdout[-800:])
        else:
            os.makedirs(os.path.join(w, "r"))
            flags = rng.choice([[], ["-l", "5000"], ["-l", "1234"], ["-l", "100000"]]) + rng.choice([[], ["-b", "40"]])
            ok = all(run([REF] + a, w, refenv).returncode == 0 for a in (["dict", "seq.mf"], ["hist", "-e", "-t", "2", "-H", "r"] + flags + ["motifs.jaspar", "seq.mf"]))
            if not ok:
                skipped += 1
                continue
            r = run([CLI, "hist", "-e", "-H", "b"] + flags + ["motifs.jaspar", "seq.mf"], w, e)
            fr, fb = sorted(os.listdir(os.path.join(w, "r"))), sorted(os.listdir(os.path.join(w, "b")))
            if r.returncode != 0 or fr != fb:
                bad += 1
                print("FAIL seed", it, r.stderr[-300:])
                continue
            for f in fr:
                if not f.endswith(".dat"):
                    continue
                a = open(os.path.join(w, "r", f)).read().split("\n")
                b = open(os.path.join(w, "b", f)).read().split("\n")
                if a == b:
                    identical += 1
                    continue
                ca = [int(x.split("\t")[1]) for x in a[1:] if x]
                cb = [int(x.split("\t")[1]) for x in b[1:] if x]
                moved = sum(abs(x - y) for x, y in zip(ca, cb)) // 2
                if a[0] == b[0] and sum(ca) == sum(cb) and moved <= max(2, sum(ca) // 1000):
                    tolerated += 1
                else:
                    bad += 1
                    print("MISMATCH seed", it, f, a[0], b[0], sum(ca), sum(cb), moved)
    shutil.rmtree(base, ignore_errors=True)
    print("%s: %d runs from seed %d%s -- identical %d, within the tolerance %d, refused by the reference %d, BAD %d"
          % (what, runs, first, " (long motifs)" if long_motifs else "", identical, tolerated, skipped, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
