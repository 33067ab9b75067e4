"""The blamm-b200 command line's HOST logic on the CPU suite (no GPU): `scan` runs against tests/mock/mock_b200scan.cpp, a
test-only stand-in for libb200scan.so that answers the C ABI from the CPU oracle and can pretend to be several devices, finish
out of order and refuse blocks as too dense.  What is under test is everything around the kernels: the reader and packer, one
worker per device with three slots, group changes, chunks dealt to devices, halving of refused chunks, the formatter and the
stream-order emission -- against the UNMODIFIED reference binary (oracle/_ref/blamm) where it is built, else the oracle's own
`scan` restatement.  The CUDA path itself is the business of tests/test_gpu_parity.py; the product library has no CPU path
(tests/test_host.py::test_create_fails_loudly_without_gpu) and nothing outside tests/ can load this stand-in."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from blamm_b200 import lib_dir, synth
from oracle import oracle as O
from tests import util

ROOT = util.ROOT
CLI = os.path.join(lib_dir(), "blamm-b200")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "blamm")
MOCK_DIR = os.path.join(ROOT, "tests", "mock", "_build")


@pytest.fixture(scope="module")
def mock_env():
    """Build the stand-in (g++, linked against oracle/liboracle.so) and return the environment that makes the CLI load it: the
    CLI finds libb200scan.so through DT_RUNPATH=$ORIGIN, which LD_LIBRARY_PATH precedes."""
    O.lib()                                                   # builds oracle/liboracle.so if it is missing
    os.makedirs(MOCK_DIR, exist_ok=True)
    so = os.path.join(MOCK_DIR, "libb200scan.so")
    src = os.path.join(ROOT, "tests", "mock", "mock_b200scan.cpp")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", so, "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                               "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lpthread"])
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = MOCK_DIR + ":" + env.get("LD_LIBRARY_PATH", "")
    env.pop("BLAMM_B200_HITS", None); env.pop("BLAMM_B200_ASCII", None); env.pop("BLAMM_B200_CHUNK", None)
    return env


def _scan(work, env, *flags, **extra):
    e = dict(env, **{k: str(v) for k, v in extra.items()})
    r = subprocess.run([CLI, "scan"] + list(flags) + ["motifs.jaspar", "seq.mf"], cwd=work, env=e, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return open(os.path.join(work, "occurrences.txt"), "rb").read(), r.stdout


def _make_inputs(work, seed, n_groups=3, n_motifs=30, max_len=14):
    """several manifest groups with distinct backgrounds, several files per group, N runs, lower-case stretches, a record shorter
    than a motif; motifs at most 14 long, where the reference's sgemm sums in position order"""
    rng = np.random.default_rng(seed)
    synth.make_jaspar_like(os.path.join(work, "motifs.jaspar"), n_motifs, seed, uniform_len=(5, max_len))
    manifest = []
    for g in range(n_groups):
        gc = 0.36 + 0.07 * g
        probs = ((1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2)
        for f in range(1 + g % 2):
            seq = synth.random_acgt(150_000 + 11_003 * g, seed * 10 + 3 * g + f, probs)
            for _ in range(4):
                a = int(rng.integers(0, len(seq) - 3000)); seq[a:a + int(rng.integers(1, 2000))] = ord("N")
                b = int(rng.integers(0, len(seq) - 3000)); seq[b:b + int(rng.integers(1, 2500))] |= 0x20
            recs = [("g%df%dchr1" % (g, f), seq[:70_000]), ("g%df%dtiny" % (g, f), seq[70_000:70_009]), ("g%df%dchr2 note" % (g, f), seq[70_009:])]
            synth.write_fasta(os.path.join(work, "g%d_%d.fa" % (g, f)), recs)
            manifest.append("grp%d\tg%d_%d.fa\n" % (g, g, f))
    open(os.path.join(work, "seq.mf"), "w").write("".join(manifest))


def _reference_lines(work, flags, mode):
    """sorted occurrence lines of the reference for the inputs in `work` (its `dict` and `hist` outputs stay there for the CLI)"""
    if os.path.exists(REF_BIN):
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
        ob = os.path.join(ROOT, "oracle", "_ref", "openblas_dir.txt")
        if os.path.exists(ob):
            env["LD_LIBRARY_PATH"] = open(ob).read().strip() + ":" + env.get("LD_LIBRARY_PATH", "")
        for args in (["dict", "seq.mf"], ["hist", "motifs.jaspar", "seq.mf"], ["scan", "-t", "2", "-o", "ref_occurrences.txt"] + flags + ["motifs.jaspar", "seq.mf"]):
            r = subprocess.run([REF_BIN] + args, cwd=work, env=env, capture_output=True, text=True)
            assert r.returncode == 0, r.stdout + r.stderr
        return sorted(open(os.path.join(work, "ref_occurrences.txt")).read().splitlines(True))
    for args in (["dict", "seq.mf"], ["hist", "motifs.jaspar", "seq.mf"]):               # the CLI's own CPU modules (byte-identical to the reference's, tests/test_host.py)
        subprocess.run([CLI] + args, cwd=work, check=True, stdout=subprocess.DEVNULL)
    lines, _ = O.scan("motifs.jaspar", "seq.mf", mode[0], mode[1], mode[2], histdir=".", base_dir=str(work))
    return sorted(lines)


@pytest.mark.parametrize("seed,flags,mode", [(201, ["-rc", "-pt", "0.001"], ("pt", 0.001, True)), (202, ["-rc"], ("rt", 0.95, True)),
                                             (203, ["-at", "7.5"], ("at", 7.5, False))])
def test_cli_scan_on_mock_devices_matches_reference(tmp_path, mock_env, seed, flags, mode):
    """Three groups (five files) through `blamm-b200 scan`: the occurrence set of the reference, and THE SAME FILE byte for byte
    whether one device scores everything or five devices that finish out of order share chunks of 20,000 characters (stream-order
    merge), with the unordered 12-byte records and the host sort, and with the character hand-over."""
    work = str(tmp_path)
    _make_inputs(work, seed)
    want = _reference_lines(work, flags, mode)
    assert len(want) > 200
    one, out = _scan(work, mock_env, *flags, "-g", "1", MOCK_B200SCAN_DEVICES=1)
    assert "Using 1 GPU devices" in out
    assert sorted(one.decode().splitlines(True)) == want
    many, out = _scan(work, mock_env, *flags, "-t", "3", MOCK_B200SCAN_DEVICES=5, MOCK_B200SCAN_DELAY_US=4000, BLAMM_B200_CHUNK=20000)
    assert "Using 5 GPU devices" in out
    assert sorted(many.decode().splitlines(True)) == want
    # chunks in stream order, hits in (position, column) order inside: the file does not depend on the devices or on their timing
    again, _ = _scan(work, mock_env, *flags, "-g", "2", MOCK_B200SCAN_DEVICES=5, MOCK_B200SCAN_DELAY_US=1500, BLAMM_B200_CHUNK=20000)
    single, _ = _scan(work, mock_env, *flags, "-g", "1", MOCK_B200SCAN_DEVICES=5, BLAMM_B200_CHUNK=20000)
    assert many == again == single
    sorted12, _ = _scan(work, mock_env, *flags, MOCK_B200SCAN_DEVICES=3, MOCK_B200SCAN_DELAY_US=1500, BLAMM_B200_CHUNK=20000, BLAMM_B200_HITS=12)
    assert sorted12 == many
    chars, _ = _scan(work, mock_env, *flags, MOCK_B200SCAN_DEVICES=2, BLAMM_B200_CHUNK=20000, BLAMM_B200_ASCII=1)
    assert chars == many


def test_cli_halves_chunks_the_device_refuses(tmp_path, mock_env):
    """A device that refuses every block with more than 300 hits (B200SCAN_ENOMEM at collect): the workers score such chunks in
    halves, recursively, keep the chunks collected behind them in order, and the file is the one of an unconstrained run --
    with several devices as well."""
    work = str(tmp_path)
    _make_inputs(work, 204, n_groups=2)
    flags = ["-rc", "-pt", "0.002"]
    want = _reference_lines(work, flags, ("pt", 0.002, True))
    free, _ = _scan(work, mock_env, *flags, BLAMM_B200_CHUNK=30000)
    assert sorted(free.decode().splitlines(True)) == want and len(want) > 2000
    tight, _ = _scan(work, mock_env, *flags, BLAMM_B200_CHUNK=30000, MOCK_B200SCAN_HIT_BUDGET=300)
    assert tight == free
    tight3, _ = _scan(work, mock_env, *flags, BLAMM_B200_CHUNK=30000, MOCK_B200SCAN_HIT_BUDGET=300, MOCK_B200SCAN_DEVICES=3, MOCK_B200SCAN_DELAY_US=2000)
    assert tight3 == free
    # a budget no halving can meet ends the run with an error, not with a truncated file that looks complete
    r = subprocess.run([CLI, "scan"] + flags + ["motifs.jaspar", "seq.mf"], cwd=work, capture_output=True, text=True,
                       env=dict(mock_env, BLAMM_B200_CHUNK="30000", MOCK_B200SCAN_HIT_BUDGET="0"))
    assert r.returncode == 1 and "too dense" in r.stderr


def test_cli_simple_mode_on_mock_devices(tmp_path, mock_env):
    """`-s`: lower case scored like upper case (the reference's naive path, motif.cpp:138-149), here with four devices."""
    work = str(tmp_path)
    _make_inputs(work, 205, n_groups=2)
    if os.path.exists(REF_BIN):
        want = _reference_lines(work, ["-s", "-rc", "-at", "8"], None)
    else:
        for args in (["dict", "seq.mf"],):
            subprocess.run([CLI] + args, cwd=work, check=True, stdout=subprocess.DEVNULL)
        lines, _ = O.scan("motifs.jaspar", "seq.mf", "at", 8.0, True, histdir=".", base_dir=work, lower_fold=True)
        want = sorted(lines)
    got, _ = _scan(work, mock_env, "-s", "-rc", "-at", "8", MOCK_B200SCAN_DEVICES=4, MOCK_B200SCAN_DELAY_US=1000, BLAMM_B200_CHUNK=25000)
    assert len(want) > 100 and sorted(got.decode().splitlines(True)) == want
    plain, _ = _scan(work, mock_env, "-rc", "-at", "8", MOCK_B200SCAN_DEVICES=4, BLAMM_B200_CHUNK=25000)
    assert plain != got                                        # the lower-case stretches make the two rules differ


def test_cli_on_the_example_with_mock_devices(golden, tmp_path, mock_env):
    """The reference's own example (two groups) in the four threshold modes of the golden fixtures, with three devices: the golden
    occurrence files (sorted text of the compiled reference) and its PWMthresholds.txt."""
    work = tmp_path / "ex"
    shutil.copytree(os.path.join(golden, "example"), work)
    os.rename(work / "sequences.mf", work / "seq.mf")
    os.rename(work / "sequences.mf.dict", work / "seq.mf.dict")
    for mode_key, flags in (("pt_rc", ["-rc", "-pt", "0.0001"]), ("pt_fwd", ["-pt", "0.0001"]), ("rt_rc", ["-rc"]), ("at_rc", ["-rc", "-at", "9.5"])):
        got, out = _scan(str(work), mock_env, *flags, MOCK_B200SCAN_DEVICES=3, MOCK_B200SCAN_DELAY_US=500, BLAMM_B200_CHUNK=30000)
        lines = sorted(got.decode().splitlines(True))
        assert lines == open(os.path.join(golden, "example", "occ_%s.txt" % mode_key)).read().splitlines(True)
        assert ("Wrote %d matches" % len(lines)) in out
        if mode_key == "pt_rc":
            assert open(work / "PWMthresholds.txt").read() == open(os.path.join(golden, "example", "PWMthresholds_pt_rc.txt")).read()


def test_cli_smallest_chunks_and_stats_account(tmp_path, mock_env):
    """Chunks of 1,024 characters (the smallest the CLI accepts: hundreds of chunks per group, every halo and fragment rule at a
    chunk border) on three devices: still the reference's occurrence set and the same file as one large chunk; `--stats` writes a
    JSON account whose totals agree with the run."""
    import json
    work = str(tmp_path)
    _make_inputs(work, 206, n_groups=2)
    flags = ["-rc", "-pt", "0.001"]
    want = _reference_lines(work, flags, ("pt", 0.001, True))
    big, _ = _scan(work, mock_env, *flags)
    tiny, out = _scan(work, mock_env, *flags, "--stats", "stats.json", MOCK_B200SCAN_DEVICES=3, MOCK_B200SCAN_DELAY_US=300, BLAMM_B200_CHUNK=1024)
    assert sorted(big.decode().splitlines(True)) == want and tiny == big
    st = json.load(open(os.path.join(work, "stats.json")))
    assert st["matches"] == len(want) and st["columns"] == 60 and len(st["devices"]) == 3
    assert sum(d["hits"] for d in st["devices"]) == len(want)
    valid = sum(int(l.split("\t")[1]) for l in open(os.path.join(work, "seq.mf.dict")) if l.startswith("TOT_SEQ_LENGTH"))
    assert sum(d["characters"] for d in st["devices"]) == valid and sum(d["chunks"] for d in st["devices"]) >= valid // 1024
    assert ("Wrote %d matches" % len(want)) in out


def test_cli_stale_dictionary_and_empty_group_on_mock(tmp_path, mock_env):
    """A group whose FASTA file holds no valid character yields nothing and does not disturb its neighbours; a FASTA file that grew a
    record after `dict` ends the run with an error and exit code 1 (the formatting threads meet a record without a name)."""
    work = str(tmp_path)
    synth.make_jaspar_like(os.path.join(work, "motifs.jaspar"), 8, 5, uniform_len=(6, 9))
    seq = synth.random_acgt(90_000, 21)
    synth.write_fasta(os.path.join(work, "a.fa"), [("s1", seq[:45_000]), ("s2", seq[45_000:])])
    open(os.path.join(work, "n.fa"), "w").write(">only_gaps\n" + "N" * 300 + "\n")
    synth.write_fasta(os.path.join(work, "b.fa"), [("t1", seq[10_000:50_000])])
    open(os.path.join(work, "seq.mf"), "w").write("g1\ta.fa\ngap\tn.fa\ng2\tb.fa\n")
    subprocess.run([CLI, "dict", "seq.mf"], cwd=work, check=True, stdout=subprocess.DEVNULL)
    got, out = _scan(work, mock_env, "-rc", "-at", "2", MOCK_B200SCAN_DEVICES=2, BLAMM_B200_CHUNK=7000)
    lines, _ = O.scan("motifs.jaspar", "seq.mf", "at", 2.0, True, histdir=".", base_dir=work)
    assert len(lines) > 100 and sorted(got.decode().splitlines(True)) == sorted(lines)
    assert not any(l.startswith("only_gaps") for l in got.decode().splitlines())
    synth.write_fasta(os.path.join(work, "a.fa"), [("s1", seq[:30_000]), ("s2", seq[30_000:60_000]), ("s3", seq[60_000:])])
    r = subprocess.run([CLI, "scan", "-rc", "-at", "2", "motifs.jaspar", "seq.mf"], cwd=work, capture_output=True, text=True,
                       env=dict(mock_env, MOCK_B200SCAN_DEVICES="2", BLAMM_B200_CHUNK="7000"))
    assert r.returncode == 1, (r.returncode, r.stderr[-500:])
    assert r.stderr.strip() != "" and "bye" not in r.stdout


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/blamm (the compiled reference) has not been built")
def test_cli_empirical_histograms_on_mock_devices_match_reference(tmp_path, mock_env):
    """`hist -e -g N`: the chunks of a group are dealt to all devices, every device keeps its own bin counters and the host adds
    them up (hist.cpp:95-160 runs histThread on -t host threads).  With four stand-in devices and chunks of 9,000 characters every
    histogram file must equal the reference binary's `hist -e` byte for byte -- and the one-device, one-chunk run's."""
    work = str(tmp_path)
    _make_inputs(work, 207, n_groups=2, n_motifs=20)
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    ob = os.path.join(ROOT, "oracle", "_ref", "openblas_dir.txt")
    if os.path.exists(ob):
        env["LD_LIBRARY_PATH"] = open(ob).read().strip() + ":" + env.get("LD_LIBRARY_PATH", "")
    os.makedirs(os.path.join(work, "ref"))
    for args in (["dict", "seq.mf"], ["hist", "-e", "-t", "3", "-H", "ref", "motifs.jaspar", "seq.mf"]):
        r = subprocess.run([REF_BIN] + args, cwd=work, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    want = {f: open(os.path.join(work, "ref", f), "rb").read() for f in os.listdir(os.path.join(work, "ref"))}
    assert len(want) == 2 * 2 * 20
    for name, extra in (("one", {}), ("four", {"MOCK_B200SCAN_DEVICES": "4", "BLAMM_B200_CHUNK": "9000"}),
                        ("two_of_four", {"MOCK_B200SCAN_DEVICES": "4", "BLAMM_B200_CHUNK": "20011"})):
        os.makedirs(os.path.join(work, name))
        flags = ["-g", "2"] if name == "two_of_four" else []
        r = subprocess.run([CLI, "hist", "-e", "-H", name] + flags + ["motifs.jaspar", "seq.mf"], cwd=work, env=dict(mock_env, **extra), capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-1000:]
        got = {f: open(os.path.join(work, name, f), "rb").read() for f in os.listdir(os.path.join(work, name))}
        assert sorted(got) == sorted(want)
        for f in want:
            assert got[f] == want[f], (name, f)


def _odd_fasta(rng):
    """records with odd headers (blanks, tabs, long names, descriptions), empty lines, lines of 0 to 3000 characters, lower case,
    N and other IUPAC letters, LF or CRLF, with or without a final line end"""
    out = []
    for r in range(rng.randint(1, 4)):
        out.append(">" + rng.choice(["s", "chr", "x y", "tab\tsep", "a" * rng.randint(1, 30)]) + str(r) + rng.choice(["", " desc"]))
        for _ in range(rng.randint(0, 40)):
            n = rng.choice([0, 1, 5, 60, 61, 200, rng.randint(1, 3000)])
            alpha = rng.choice(["ACGT", "ACGT", "ACGTacgt", "ACGTN", "ACGTacgtNnRY"])
            out.append("".join(rng.choice(alpha) for _ in range(n)))
    eol = rng.choice(["\n", "\n", "\r\n"])
    return eol.join(out) + (eol if rng.random() < 0.8 else "")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/blamm (the compiled reference) has not been built")
def test_cli_fuzz_against_reference_binary(tmp_path, mock_env):
    """Seeded fuzz of the whole command line against the UNMODIFIED reference: odd FASTA files x small random motif sets x the
    threshold modes, with random device counts, chunk sizes, hand-overs and record formats on the stand-in library.  `dict` must be
    byte-identical; `scan` must give the same occurrence text, or -- where the reference's sgemm sums in another order (it does for
    some matrix shapes even with short motifs: a handful of scores then differ in the sixth digit) -- pass north_star's rule as
    tools/parity_list.py implements it: identical sets, scores within 1e-4, exceptions only within 1e-4 of their threshold."""
    import random
    import sys
    refenv = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    ob = os.path.join(ROOT, "oracle", "_ref", "openblas_dir.txt")
    if os.path.exists(ob):
        refenv["LD_LIBRARY_PATH"] = open(ob).read().strip() + ":" + refenv.get("LD_LIBRARY_PATH", "")
    modes = [["-pt", "0.01"], ["-rt", "0.8"], ["-at", "3"], ["-rc", "-pt", "0.005"], ["-rc", "-at", "4.5"], ["-s", "-rc", "-at", "5"]]
    identical = within_tolerance = 0
    for it in range(14):
        rng = random.Random(5000 + it)
        w = tmp_path / ("it%d" % it)
        os.makedirs(w / "b2")
        synth.make_jaspar_like(str(w / "motifs.jaspar"), rng.randint(1, 12), 5000 + it, uniform_len=(5, 14))
        manifest = []
        for f in range(rng.randint(1, 3)):
            open(w / ("f%d.fa" % f), "w", newline="").write(_odd_fasta(rng))
            manifest.append("g%d\tf%d.fa\n" % (f % 2, f))
        open(w / "seq.mf", "w").write("".join(manifest))
        mode = rng.choice(modes)
        ok = True
        for args in (["dict", "seq.mf"], ["hist", "motifs.jaspar", "seq.mf"], ["scan", "-t", "2", "-o", "ref.txt"] + mode + ["motifs.jaspar", "seq.mf"]):
            ok = ok and subprocess.run([REF_BIN] + args, cwd=w, env=refenv, capture_output=True, text=True).returncode == 0
        # dict on its own copy of the inputs, with 1 to 5 parser threads
        for f in os.listdir(w):
            if f.endswith(".fa") or f == "seq.mf":
                shutil.copy(w / f, w / "b2" / f)
        r = subprocess.run([CLI, "dict", "seq.mf"], cwd=w / "b2", env=dict(os.environ, BLAMM_B200_INGEST_THREADS=str(rng.choice([1, 2, 5]))),
                           capture_output=True, text=True)
        if not ok:
            continue                                      # (the reference refused the input)
        assert r.returncode == 0 and (w / "b2" / "seq.mf.dict").read_bytes() == (w / "seq.mf.dict").read_bytes(), it
        extra = {"MOCK_B200SCAN_DEVICES": rng.randint(1, 4), "BLAMM_B200_CHUNK": rng.choice([1024, 3000, 50000, 1 << 25]),
                 "MOCK_B200SCAN_DELAY_US": rng.choice([0, 500])}
        if rng.random() < 0.3:
            extra["BLAMM_B200_HITS"] = 12
        if rng.random() < 0.3:
            extra["BLAMM_B200_ASCII"] = 1
        got, _ = _scan(str(w), mock_env, "-o", "occurrences.txt", *mode, **extra)
        if sorted(got.decode().splitlines(True)) == sorted(open(w / "ref.txt").read().splitlines(True)):
            identical += 1
            continue
        m = [x for x in mode if x != "-s"]
        cmd = [sys.executable, os.path.join(ROOT, "tools", "parity_list.py"), "--ours", str(w / "occurrences.txt"), "--ref", str(w / "ref.txt"),
               "--motifs", str(w / "motifs.jaspar"), "--manifest", str(w / "seq.mf"), "--histdir", str(w)]
        if "-rc" in m:
            cmd.append("--rc"); m.remove("-rc")
        q = subprocess.run(cmd + ["--" + m[0][1:], m[1]], capture_output=True, text=True)
        assert q.returncode == 0 and "PARITY OK" in q.stdout, (it, mode, q.stdout[-1500:])
        within_tolerance += 1
    assert identical >= 8 and identical + within_tolerance >= 12


@pytest.mark.parametrize("devices", [1, 3])
def test_cli_survives_injected_device_failures(tmp_path, mock_env, devices):
    """Fault injection (the reference ignores every CUDA return code but cudaMalloc's): the n-th submit, collect
    or set_motifs of the run fails with B200SCAN_ECUDA, early, in the middle of a group and near the end, with one and with three
    devices and chunks in flight on all of them.  Every such run must END -- no worker, reader or emitter left waiting for a turn
    that cannot come -- with exit code 1 and the library's error text, and must not claim success."""
    work = str(tmp_path)
    _make_inputs(work, 208, n_groups=2)
    for args in (["dict", "seq.mf"], ["hist", "motifs.jaspar", "seq.mf"]):
        subprocess.run([CLI] + args, cwd=work, check=True, stdout=subprocess.DEVNULL)
    base = dict(mock_env, MOCK_B200SCAN_DEVICES=str(devices), BLAMM_B200_CHUNK="9000", MOCK_B200SCAN_DELAY_US="300")
    ok = subprocess.run([CLI, "scan", "-rc", "-pt", "0.001", "motifs.jaspar", "seq.mf"], cwd=work, env=base, capture_output=True, text=True, timeout=120)
    assert ok.returncode == 0 and "Wrote" in ok.stdout
    for var, counts in (("MOCK_B200SCAN_FAIL_SUBMIT", (1, 2, 17, 40)), ("MOCK_B200SCAN_FAIL_COLLECT", (1, 3, 18, 41)), ("MOCK_B200SCAN_FAIL_MOTIFS", (1, devices + 1))):
        for n in counts:
            r = subprocess.run([CLI, "scan", "-rc", "-pt", "0.001", "motifs.jaspar", "seq.mf"], cwd=work, env=dict(base, **{var: str(n)}),
                               capture_output=True, text=True, timeout=120)
            assert r.returncode == 1, (var, n, r.returncode, r.stdout[-300:], r.stderr[-300:])
            assert "injected failure" in r.stderr and "Wrote" not in r.stdout and "bye" not in r.stdout, (var, n, r.stderr[-300:])


def test_cli_hist_e_survives_injected_device_failures(tmp_path, mock_env):
    """The same for `hist -e`: a histogram block or a set_motifs that fails on one of three devices ends the run with exit code 1."""
    work = str(tmp_path)
    _make_inputs(work, 209, n_groups=2, n_motifs=12)
    subprocess.run([CLI, "dict", "seq.mf"], cwd=work, check=True, stdout=subprocess.DEVNULL)
    os.makedirs(os.path.join(work, "he"))
    base = dict(mock_env, MOCK_B200SCAN_DEVICES="3", BLAMM_B200_CHUNK="9000")
    ok = subprocess.run([CLI, "hist", "-e", "-H", "he", "motifs.jaspar", "seq.mf"], cwd=work, env=base, capture_output=True, text=True, timeout=120)
    assert ok.returncode == 0 and "bye" in ok.stdout
    for var, counts in (("MOCK_B200SCAN_FAIL_HIST", (1, 5, 30)), ("MOCK_B200SCAN_FAIL_MOTIFS", (2, 4))):
        for n in counts:
            r = subprocess.run([CLI, "hist", "-e", "-H", "he", "motifs.jaspar", "seq.mf"], cwd=work, env=dict(base, **{var: str(n)}),
                               capture_output=True, text=True, timeout=120)
            assert r.returncode == 1 and "injected failure" in r.stderr and "bye" not in r.stdout, (var, n, r.returncode, r.stderr[-300:])
