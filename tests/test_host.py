"""Host-side model (C++ through the ctypes C ABI and the blamm-b200 command line) against the golden
fixtures produced by the compiled reference and against the oracle.  No GPU, no compute calls into
libb200scan.so -- only that it loads and exports what include/b200scan.h declares."""
import ctypes
import os
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

from blamm_b200 import capi, lib_dir, synth
from oracle import oracle as O
from oracle.refdump_io import read_refdump
from tests import util

ROOT = util.ROOT
CLI = os.path.join(lib_dir(), "blamm-b200")


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:b200scan|blamm)_[a-z0-9_]+)\s*\(", txt)))


def test_scan_library_exports_every_declared_symbol():
    L = capi.scan_lib()
    names = _declared("b200scan.h")
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    assert L.b200scan_abi_version() == 2 and "b200scan_collect8" in names


def test_host_library_exports_every_declared_symbol():
    L = capi.host_lib()
    names = _declared("blamm_host.h")
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert capi.scan_lib().b200scan_device_count() == 0
    with pytest.raises(capi.ScanError) as e:
        capi.Scanner(0, 1 << 20, 1 << 10)
    assert e.value.code == -3 and "no CPU path" in str(e.value)


@pytest.mark.parametrize("case,mode_key", [("example", "pt_rc"), ("example", "pt_fwd"), ("example", "rt_rc"), ("example", "at_rc"),
                                           ("edge", "pt_rc"), ("edge", "rt_rc"), ("edge", "at_low")])
def test_matrix_and_thresholds_match_reference(golden, case, mode_key):
    d = os.path.join(golden, case)
    mode, value, rc = util.MODES[mode_key]
    ms = capi.MotifSet(os.path.join(d, "motifs.jaspar"), rc)
    dump = read_refdump(os.path.join(d, "refdump_%s.bin" % mode_key))
    for sp, r in zip(O.load_dict(os.path.join(d, "sequences.mf.dict")), dump):
        P, col_len, is_rc = ms.generate_matrix(sp.counts)
        thr = ms.thresholds(mode, value, sp.name, d)
        # same std::sort, same comparator -> identical column order as the reference, not just the same set
        assert ms.names == [c["name"] for c in r["cols"]]
        assert is_rc.tolist() == [c["rc"] for c in r["cols"]] and col_len.tolist() == [c["len"] for c in r["cols"]]
        assert np.array_equal(P.view(np.uint32), r["P"].view(np.uint32))
        assert np.array_equal(thr.view(np.uint32), np.array([c["thr"] for c in r["cols"]], dtype=np.float32).view(np.uint32))


@pytest.mark.parametrize("case", ["example", "edge"])
@pytest.mark.parametrize("payload,halo", [(1 << 30, 0), (1000, 22), (37, 5), (1, 0)])
def test_fasta_stream_matches_oracle(golden, case, payload, halo):
    d = os.path.join(golden, case)
    for sp in O.load_dict(os.path.join(d, "sequences.mf.dict")):
        want = O.build_stream(sp.files, sp.tot_len, d)
        fs = capi.FastaStream([os.path.join(d, f) for f in sp.files], sp.tot_len)
        chars, frag = bytearray(), []
        pos = 0
        while True:
            c = fs.next(payload, halo)
            if c is None:
                break
            assert c["stream_start"] == pos and c["n_payload"] <= payload and c["n_total"] <= payload + halo
            assert c["frag_start"][0] == 0
            # chunk-relative fragment table -> global (record, position) of every fragment start inside the payload
            for s, q, p in zip(c["frag_start"], c["frag_seq"], c["frag_pos"]):
                if s < c["n_payload"]:
                    frag.append((pos + int(s), int(q), int(p)))
            chars += c["chars"][:c["n_payload"]]
            # the halo must be the head of the next chunk
            halo_chars = c["chars"][c["n_payload"]:]
            assert want.chars[pos + c["n_payload"]:pos + c["n_total"]] == halo_chars
            pos += c["n_payload"]
        assert bytes(chars) == want.chars
        assert fs.seq_names()[:len(sp.seq_names)] == sp.seq_names
        # every true fragment start is reported with the same coordinates; chunk heads add redundant entries only
        true = set(zip(want.frag_start.tolist(), want.frag_seq.tolist(), want.frag_pos.tolist()))
        assert true <= set(frag)
        for s, q, p in frag:
            wq, wp = O.stream_to_seq(want, np.array([s], dtype=np.uint64))
            assert (int(wq[0]), int(wp[0])) == (q, p)


def _collect_stream(fs, payload, halo):
    """Whole stream of a FastaStream: payload characters and every fragment the chunks report, in global coordinates."""
    chars, frag, pos = bytearray(), set(), 0
    while True:
        c = fs.next(payload, halo)
        if c is None:
            break
        assert c["stream_start"] == pos and c["frag_start"][0] == 0
        for s, q, p in zip(c["frag_start"], c["frag_seq"], c["frag_pos"]):
            if s < c["n_payload"]:
                frag.add((pos + int(s), int(q), int(p)))
        chars += c["chars"][:c["n_payload"]]
        pos += c["n_payload"]
    return bytes(chars), frag


def _check_against_oracle(files, max_filtered, threads, segment_bytes, payload, halo):
    want = O.build_stream(files, max_filtered)
    fs = capi.FastaStream(files, max_filtered, threads=threads, segment_bytes=segment_bytes)
    chars, frag = _collect_stream(fs, payload, halo)
    assert chars == want.chars
    # the reference reads header lines lazily (sequence.cpp:216-228): past the length cut none is registered, while the
    # oracle lists every header of the files
    names = fs.seq_names()
    assert names == want.seq_names[:len(names)] and (len(names) == len(want.seq_names) or len(chars) == max_filtered)
    assert len(names) > int(want.frag_seq.max(initial=0)) or not len(chars)
    counts = [chars.upper().count(c) for c in b"ACGT"]
    assert fs.counts() == counts
    true = set(zip(want.frag_start.tolist(), want.frag_seq.tolist(), want.frag_pos.tolist()))
    assert true <= frag
    extra = sorted(frag - true)          # chunk heads only: same coordinates as the oracle's mapping
    if extra:
        wq, wp = O.stream_to_seq(want, np.array([e[0] for e in extra], dtype=np.uint64))
        assert [(int(a), int(b)) for a, b in zip(wq, wp)] == [(e[1], e[2]) for e in extra]


@pytest.mark.parametrize("case", ["example", "edge"])
@pytest.mark.parametrize("threads,segment_bytes", [(4, 7), (3, 64), (8, 1000), (5, 0)])
def test_parallel_ingest_matches_oracle_on_golden_inputs(golden, case, threads, segment_bytes):
    """The wave parser (segments parsed blind, carries resolved by the stitch) gives the reference's stream
    (FastaBatch, sequence.cpp:142-293) whatever the segmentation -- including the dictionary's length cut."""
    d = os.path.join(golden, case)
    for sp in O.load_dict(os.path.join(d, "sequences.mf.dict")):
        files = [os.path.join(d, f) for f in sp.files]
        for cut in (sp.tot_len, max(sp.tot_len // 3, 1), 2 ** 62):
            _check_against_oracle(files, cut, threads, segment_bytes, 1000, 22)


def _nasty_fasta(tmp_path, seed):
    """Files that exercise every border rule: lines far longer than the 4 KB split limit, CRLF, blank lines, N runs,
    lower case, a header longer than a segment, a header at the very end, no trailing newline, a second file that
    continues the open record without a header."""
    rng = np.random.default_rng(seed)

    def seq(n):
        s = rng.choice(np.frombuffer(b"ACGTacgtNnRY-", dtype=np.uint8), size=n,
                       p=[.2, .2, .2, .2, .03, .03, .03, .03, .02, .02, .01, .01, .02]).tobytes()
        return s

    f1 = b">chr1 first record\n" + seq(30_000) + b"\n" + seq(100) + b"\r\n\n\n" + seq(5000) + b"\n"
    f1 += b">" + b"h" * 3000 + b" " + b"x" * 9000 + b"\n" + b"N" * 7000 + b"\n" + seq(12_345) + b"\n>empty\n>chr3\n"
    for _ in range(200):
        f1 += seq(int(rng.integers(0, 80))) + b"\n"
    f1 += seq(20_000)                                   # no trailing newline
    f2 = seq(777) + b"\n>chr4\tdescription\n" + seq(9000) + b"\n>last"
    a, b = tmp_path / "a.fa", tmp_path / "b.fa"
    a.write_bytes(f1)
    b.write_bytes(f2)
    return [str(a), str(b)]


@pytest.mark.parametrize("threads,segment_bytes", [(1, 0), (2, 1), (4, 13), (4, 4097), (7, 5000), (8, 0), (3, 1 << 16)])
def test_parallel_ingest_border_rules(tmp_path, threads, segment_bytes):
    files = _nasty_fasta(tmp_path, 11)
    total = len(O.build_stream(files).chars)
    for cut in (2 ** 62, total, total - 1, 30_001, 29_999, 1):
        _check_against_oracle(files, cut, threads, segment_bytes, 10_000, 34)


def test_parallel_ingest_is_independent_of_threads_at_size(tmp_path):
    """4 MB of 60-column FASTA with N runs: 1, 3 and 8 threads with request-sized segments give identical chunks."""
    p = tmp_path / "g.fa"
    seq = synth.random_acgt(4_000_000, 5)
    for a in (100_000, 1_234_567, 3_999_000):
        seq[a:a + 777] = ord("N")
    synth.write_fasta(str(p), [("chrA", seq[:1_500_000]), ("chrB desc", seq[1_500_000:1_500_050]), ("chrC", seq[1_500_050:])])
    ref = None
    for threads in (1, 3, 8):
        fs = capi.FastaStream([str(p)], threads=threads)
        got = _collect_stream(fs, 700_000, 34) + (fs.seq_names(), fs.counts())
        ref = ref or got
        assert got == ref
    assert ref[0] == O.build_stream([str(p)]).chars


def _pack_restated(buf, fold_lower):
    """The packing rule of include/blamm_host.h / csrc/pack.cuh, one character at a time."""
    n = len(buf)
    code = np.zeros(256, dtype=np.uint32); zero = np.ones(256, dtype=np.uint32)
    for k, (u, l) in enumerate(zip(b"ACGT", b"acgt")):
        code[u] = code[l] = k
        zero[u] = 0
        zero[l] = 0 if fold_lower else 1
    m = (n + 31) // 32 * 32
    c = np.zeros(m, dtype=np.uint32); z = np.ones(m, dtype=np.uint32)
    c[:n] = code[buf]; z[:n] = zero[buf]
    codes = (c.reshape(-1, 16) << (2 * np.arange(16, dtype=np.uint32))).sum(axis=1, dtype=np.uint64).astype(np.uint32)[:(n + 15) // 16]
    zm = (z.reshape(-1, 32) << np.arange(32, dtype=np.uint32)).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    return codes, zm, bool(z[:n].any())


@pytest.mark.parametrize("lower", [capi.LOWER_ZERO, capi.LOWER_FOLD])
def test_host_packer_matches_its_rule(lower):
    """blamm_pack_ascii (the host twin of the device packer) on every length around the word sizes, mixed case, and
    bytes that a filtered stream never holds (every byte value occurs)."""
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"ACGTacgt", dtype=np.uint8)
    for n in [0, 1, 7, 8, 9, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 1000, 100_003]:
        buf = alphabet[rng.integers(0, 8, size=n)].copy()
        codes, zm, hz = capi.pack_ascii(buf, lower)
        wc, wz, wh = _pack_restated(buf, lower == capi.LOWER_FOLD)
        assert np.array_equal(codes, wc) and np.array_equal(zm, wz) and hz == wh, n
        up = alphabet[rng.integers(0, 4, size=n)].copy()          # upper case only: no character contributes zero
        assert capi.pack_ascii(up, lower)[2] is False
    every = np.tile(np.arange(256, dtype=np.uint8), 5)[:1273]
    rng.shuffle(every)
    codes, zm, hz = capi.pack_ascii(every, lower)
    wc, wz, wh = _pack_restated(every, lower == capi.LOWER_FOLD)
    assert np.array_equal(codes, wc) and np.array_equal(zm, wz) and hz and wh


@pytest.mark.parametrize("threads", [1, 5])
def test_fasta_stream_packs_its_chunks(tmp_path, threads):
    """blamm_fasta_pack: every chunk of a soft-masked FASTA (2.5 MB: several 1 MiB pack slices, ragged last chunk) packed on
    the parser threads == the rule applied to the chunk's characters, both lower-case modes."""
    p = tmp_path / "g.fa"
    seq = synth.random_acgt(2_500_000, 21)
    seq[300_000:1_200_000] |= 0x20
    seq[2_000_000:2_000_500] = ord("N")
    synth.write_fasta(str(p), [("chr1", seq[:1_700_000]), ("chr2", seq[1_700_000:])])
    fs = capi.FastaStream([str(p)], threads=threads)
    n_chunks = 0
    while True:
        c = fs.next(1_100_000, 29)
        if c is None:
            break
        buf = np.frombuffer(c["chars"], dtype=np.uint8)
        for lower in (capi.LOWER_ZERO, capi.LOWER_FOLD):
            codes, zm, hz = fs.pack(c["n_total"], lower)
            wc, wz, wh = _pack_restated(buf, lower == capi.LOWER_FOLD)
            assert np.array_equal(codes, wc) and np.array_equal(zm, wz) and hz == wh
        n_chunks += 1
    assert n_chunks == 3


def test_format_score_matches_printf():
    """blamm_format_score (fast "%g" of the occurrence writer, pwmscan.cpp:94) == C printf("%g") on random bit patterns of
    every magnitude, on scores as a scan produces them, and on the edges of its fast path (ties, carries into the next
    decade, the switch to exponent notation).  tools/format_check.cpp runs the same comparison over every float in
    [1e-5, 1e7]."""
    L = capi.host_lib()
    buf = ctypes.create_string_buffer(40)

    def fmt(v):
        n = L.blamm_format_score(ctypes.c_float(float(v)), buf)
        return buf.raw[:n].decode()

    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2 ** 32, size=60_000, dtype=np.uint64).astype(np.uint32)
    vals = bits.view(np.float32)
    vals = vals[~np.isnan(vals)]                                  # (printf prints the sign of a NaN, Python does not)
    scores = rng.uniform(-40, 40, size=60_000).astype(np.float32)
    small = (rng.uniform(-1, 1, size=20_000) * 10.0 ** rng.integers(-6, 8, size=20_000)).astype(np.float32)
    edges = np.array([0.0, -0.0, 1.0, 9.999995, 9.9999949, 99999.95, 999999.4, 999999.5, 999999.94, 1e6, 100000.5, 100001.5, 1e-4,
                      9.9999e-5, 0.00099999994, 0.001, 0.01, 0.1, 0.5, 123456.7, 2.5e-4, 1.0000005, 7.25, -12.0625, 3.4e38, 1e-45,
                      np.inf, -np.inf], dtype=np.float32)
    for v in np.concatenate([vals, scores, small, edges]):
        assert fmt(v) == "%g" % float(v), (float(v), fmt(v))


def test_format_score_sweep_against_printf(tmp_path):
    """tools/format_check.cpp: every 61st float between 1e-5 and 1e7 (both signs, 11 M values) through blamm_format_score and
    through snprintf("%g"); the full sweep (stride 1, 669 M values, 30 s on 8 cores) was run when the fast path was written."""
    exe = str(tmp_path / "format_check")
    lib = lib_dir()
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tools", "format_check.cpp"), "-I" + os.path.join(ROOT, "include"),
                    "-L" + lib, "-lblammhost", "-Wl,-rpath," + lib, "-lpthread", "-o", exe], check=True)
    r = subprocess.run([exe, "61"], capture_output=True, text=True)
    assert r.returncode == 0 and " 0 mismatches" in r.stdout, r.stdout


@pytest.mark.parametrize("n_hits,threads", [(0, 2), (100, 3), (5000, 1), (300_000, 4), (1_000_000, 7)])
def test_occurrence_writer_selftest(n_hits, threads):
    """`blamm-b200 selftest-writer` (no GPU): the CLI's occurrence writer -- range partition, radix sort by (position, column),
    line formatting, in-order writer thread -- for 16-byte and 12-byte hit records, and for the ordered 8-byte records + bucket
    index the device hands over under B200SCAN_HITS_8 (no host sort), against a plain std::sort + snprintf restatement of
    the reference's line format (pwmscan.cpp:88-95), compiled into the same binary."""
    # in tmpfs the writer copies through shared mappings, elsewhere it uses pwrite: both, and both forced
    for env in ({}, {"TMPDIR": "/dev/shm"}, {"BLAMM_B200_WRITER": "mmap"}, {"TMPDIR": "/dev/shm", "BLAMM_B200_WRITER": "pwrite"}):
        if env.get("TMPDIR") and not os.path.isdir(env["TMPDIR"]):
            continue
        r = subprocess.run([CLI, "selftest-writer", str(n_hits), str(threads)], capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.count("identical") == 3 and "DIFFERENT" not in r.stdout


@pytest.mark.parametrize("chunks,workers,threads", [(40, 8, 4), (33, 2, 3), (1, 4, 1), (12, 12, 2)])
def test_stream_order_merge_selftest(chunks, workers, threads):
    """`blamm-b200 selftest-order` (no GPU): the host-side merge of `scan -g N`.  N threads stand in for the per-GPU workers and push
    chunks of ordered 8-byte records through the product's formatter and emitter after random delays; the occurrence file must be
    byte for byte what ONE worker writes -- every chunk in its place in the stream whichever worker finished first -- for chunks
    dealt round robin and for chunks taken by whoever is free; in tmpfs (shared mappings) and on disk (pwrite)."""
    for env in ({}, {"TMPDIR": "/dev/shm"}):
        if env.get("TMPDIR") and not os.path.isdir(env["TMPDIR"]):
            continue
        r = subprocess.run([CLI, "selftest-order", str(chunks), str(workers), str(threads)], capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.count("identical") == 2 and "DIFFERENT" not in r.stdout and "in stream order" in r.stdout


def test_fasta_rejects_headerless_input(tmp_path):
    p = tmp_path / "bad.fa"
    p.write_text("ACGT\n>late\nACGT\n")
    fs = capi.FastaStream([str(p)])
    with pytest.raises(capi.HostError, match="fasta format"):
        fs.next(100, 0)


def test_cli_dict_and_hist_are_byte_identical_to_reference(golden, tmp_path):
    for case in ("example", "edge"):
        src = os.path.join(golden, case)
        work = tmp_path / case
        shutil.copytree(src, work)
        for f in os.listdir(work):
            if os.path.isfile(work / f) and (f.startswith("hist_") or f.endswith(".dict")):
                os.remove(work / f)
        subprocess.run([CLI, "dict", "sequences.mf"], cwd=work, check=True, stdout=subprocess.DEVNULL)
        assert (work / "sequences.mf.dict").read_bytes() == open(os.path.join(src, "sequences.mf.dict"), "rb").read()
        subprocess.run([CLI, "hist", "motifs.jaspar", "sequences.mf"], cwd=work, check=True, stdout=subprocess.DEVNULL)
        dats = [f for f in os.listdir(src) if f.startswith("hist_") and f.endswith(".dat") and os.path.isfile(os.path.join(src, f))]
        assert dats
        for f in dats:
            assert (work / f).read_bytes() == open(os.path.join(src, f), "rb").read(), f


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "blamm")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/blamm (the compiled reference) has not been built")
def test_cli_dict_and_hist_match_live_reference_binary(tmp_path):
    """blamm-b200 dict / hist next to the unmodified reference binary on fresh seeded inputs (120 motifs of 5..35 positions, three
    groups, several files per group, N runs, lower case, CRLF-free 60-column FASTA): `.dict`, every `hist_*.dat` / `.gnu` and
    the nucleotide report byte-identical; with 8 parser threads as well."""
    synth.make_jaspar_like(str(tmp_path / "motifs.jaspar"), 120, 404)
    rng = np.random.default_rng(404)
    manifest = []
    for g, gc in enumerate((0.36, 0.44, 0.55)):
        probs = ((1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2)
        for f in range(1 + g % 2):
            seq = synth.random_acgt(300_000 + 7_001 * g, 900 + 10 * g + f, probs)
            for _ in range(6):
                a = int(rng.integers(0, len(seq) - 4000)); seq[a:a + int(rng.integers(1, 3000))] = ord("N")
                b = int(rng.integers(0, len(seq) - 4000)); seq[b:b + int(rng.integers(1, 3000))] |= 0x20
            synth.write_fasta(str(tmp_path / ("g%d_%d.fa" % (g, f))), [("g%d_%d_a" % (g, f), seq[:100_000]), ("g%d_%d_b x" % (g, f), seq[100_000:])])
            manifest.append("grp%d\tg%d_%d.fa\n" % (g, g, f))
    open(tmp_path / "seq.mf", "w").write("".join(manifest))
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    ob = os.path.join(ROOT, "oracle", "_ref", "openblas_dir.txt")
    if os.path.exists(ob):
        env["LD_LIBRARY_PATH"] = open(ob).read().strip() + ":" + env.get("LD_LIBRARY_PATH", "")
    outs = {}
    for who, exe, extra in (("ref", REF_BIN, {}), ("b200", CLI, {"BLAMM_B200_INGEST_THREADS": "8"})):
        d = tmp_path / who
        os.makedirs(d)
        for f in os.listdir(tmp_path):
            if os.path.isfile(tmp_path / f):
                shutil.copy(tmp_path / f, d / f)
        for args in (["dict", "seq.mf"], ["hist", "motifs.jaspar", "seq.mf"]):
            r = subprocess.run([exe] + args, cwd=d, env=dict(env, **extra), capture_output=True, text=True)
            assert r.returncode == 0, r.stdout + r.stderr
        outs[who] = {f: (d / f).read_bytes() for f in os.listdir(d) if f.endswith(".dict") or f.startswith("hist_")}
    assert len(outs["ref"]) == 1 + 2 * 3 * 120 and sorted(outs["ref"]) == sorted(outs["b200"])
    for f, data in outs["ref"].items():
        assert outs["b200"][f] == data, f


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/blamm (the compiled reference) has not been built")
def test_jaspar_parser_quirks_match_live_reference_binary(tmp_path):
    """The JASPAR grammar is whatever the reference's stream extraction accepts (motif.cpp:377-407): blank lines between records,
    text after the name, tabs, a bracket glued to the last count ("12]"), signed counts ("+7"), CRLF line ends, rows that are not
    labelled A/C/G/T, a last line without a newline, and a lone header token at the very end of the file (no record).  Both
    binaries run `hist` on such a file; names, lengths and every histogram file must agree byte for byte."""
    text = (
        "\n\n>MA0001.1 first motif, description ignored\n"
        "A  [ 10  2  3 40 12 ]\n"
        "C  [  5 30  3  1  2 ]\n"
        "G  [  5  3 30  1  2 ]\n"
        "T  [  5  2  3  1 12]\n"
        "\n   \n"
        ">MA0002.2\tTAB\tseparated\r\n"
        "A\t[\t1\t+7\t9\t]\r\n"
        "C\t[\t8\t1\t0\t]\r\n"
        "G\t[\t0\t1\t0\t]\r\n"
        "T\t[\t0\t0\t0\t]\r\n"
        ">shortest\n"
        "x y 3 1 1 1 9 1 1\n"
        "x y 1 3 1 1 1 9 1\n"
        "x y 1 1 3 1 1 1 9\n"
        "x y 9 9 9 3 1 1 1 junk 5 5\n"
        "Xlast_motif trailing\n"
        "A [ 100 0 ]\nC [ 0 100 ]\nG [ 0 0 ]\nT [ 0 0 ]"
        "\n>dangling")
    open(tmp_path / "motifs.jaspar", "w", newline="").write(text)
    seq = synth.random_acgt(20_000, 5)
    synth.write_fasta(str(tmp_path / "a.fa"), [("r1", seq)])
    open(tmp_path / "seq.mf", "w").write("grp\ta.fa\n")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    ob = os.path.join(ROOT, "oracle", "_ref", "openblas_dir.txt")
    if os.path.exists(ob):
        env["LD_LIBRARY_PATH"] = open(ob).read().strip() + ":" + env.get("LD_LIBRARY_PATH", "")
    outs = {}
    for who, exe in (("ref", REF_BIN), ("b200", CLI)):
        d = tmp_path / who
        os.makedirs(d)
        for f in ("motifs.jaspar", "a.fa", "seq.mf"):
            shutil.copy(tmp_path / f, d / f)
        for args in (["dict", "seq.mf"], ["hist", "motifs.jaspar", "seq.mf"]):
            r = subprocess.run([exe] + args, cwd=d, env=env, capture_output=True, text=True)
            assert r.returncode == 0, r.stdout + r.stderr
        assert "Loaded 4 motifs" in r.stdout and "Maximum motif size: 7" in r.stdout, r.stdout
        outs[who] = {f: (d / f).read_bytes() for f in os.listdir(d) if f.startswith("hist_")}
    assert sorted(outs["ref"]) == sorted("hist_grp_%s.%s" % (n, e) for n in ("MA0001.1", "MA0002.2", "shortest", "last_motif") for e in ("dat", "gnu"))
    assert outs["ref"] == outs["b200"]


def test_cli_error_behaviour(golden, tmp_path):
    work = tmp_path / "ex"
    shutil.copytree(os.path.join(golden, "example"), work)

    def run(*args):
        return subprocess.run([CLI] + list(args), cwd=work, capture_output=True, text=True)

    assert run().returncode == 1
    assert run("scan", "-at", "1", "-rt", "0.5", "motifs.jaspar", "sequences.mf").returncode == 1
    r = run("scan", "-rt", "1.5", "motifs.jaspar", "sequences.mf")
    assert r.returncode == 1 and "range [0..1]" in r.stderr
    r = run("scan", "-pt", "0.0001", "-H", "nowhere", "motifs.jaspar", "sequences.mf")
    assert r.returncode == 1           # no GPU here, or no histogram there: either way a loud failure, never a silent CPU scan
    r = run("scan", "motifs.jaspar", "missing.mf")
    assert r.returncode == 1 and "Cannot open file" in r.stderr
    r = run("hist", "-e", "motifs.jaspar", "sequences.mf")
    import torch
    if not torch.cuda.is_available():      # the empirical mode scores on the GPU; without one it must fail loudly
        assert r.returncode == 1 and "no sm_100 devices" in r.stderr
    assert run("--version").returncode == 0


def test_synthetic_generators_are_seeded(tmp_path):
    a = synth.random_acgt(10000, 5); b = synth.random_acgt(10000, 5)
    assert np.array_equal(a, b) and set(np.unique(a)) <= set(b"ACGT")
    p1 = synth.make_jaspar_like(str(tmp_path / "a.jaspar"), 30, 9)
    p2 = synth.make_jaspar_like(str(tmp_path / "b.jaspar"), 30, 9)
    assert (tmp_path / "a.jaspar").read_text() == (tmp_path / "b.jaspar").read_text()
    assert [len(x) for x in p1][:3] == [35, 30, 5]
    motifs = O.load_jaspar(str(tmp_path / "a.jaspar"))
    assert sorted(len(m) for m in motifs) == sorted(len(x) for x in p2)


def test_parity_lister_lists_only_near_threshold_differences(tmp_path, golden):
    """tools/parity_list.py (north_star's correctness clause): identical files pass; a missing occurrence FAR from its threshold
    fails; an occurrence present on one side only whose score lies within 1e-4 of its threshold is listed and tolerated; a score
    that differs by more than the tolerance fails."""
    d = os.path.join(golden, "example")
    ref = open(os.path.join(d, "occ_pt_rc.txt")).read().splitlines(True)
    ms = capi.MotifSet(os.path.join(d, "motifs.jaspar"), revcompl=True)
    sp = O.load_dict(os.path.join(d, "sequences.mf.dict"))[0]
    P, col_len, is_rc = ms.generate_matrix(sp.counts)
    thr = ms.thresholds("pt", 1e-4, sp.name, d)

    def run(ours_lines, ref_lines):
        a, b = tmp_path / "a.txt", tmp_path / "b.txt"
        a.write_text("".join(ours_lines)); b.write_text("".join(ref_lines))
        return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "parity_list.py"), "--ours", str(a), "--ref", str(b),
                               "--motifs", os.path.join(d, "motifs.jaspar"), "--manifest", os.path.join(d, "sequences.mf"), "--pt", "0.0001",
                               "--rc", "--out", str(tmp_path / "list.txt")], capture_output=True, text=True)
    shuffled = ref[::-1]
    r = run(shuffled, ref)
    assert r.returncode == 0 and "identical occurrence sets" in r.stdout, r.stdout + r.stderr
    r = run(ref[1:], ref)                                             # 10.8353 is not within 1e-4 of its threshold
    assert r.returncode == 1 and "only-ref" in r.stdout and "FAILED" in r.stdout
    # a made-up occurrence 4e-5 above its threshold, on the reference side only: listed, tolerated
    c = 0
    near = "%s\tblamm\t%s\t7\t%d\t%s\t%s\t.\t.\n" % (sp.seq_names[0], ms.names[c], 7 + int(col_len[c]), O.fmt_g(np.float32(thr[c] + 4e-5)),
                                                    "-" if is_rc[c] else "+")
    r = run(ref, ref + [near])
    assert r.returncode == 0 and "within the tolerance" in r.stdout and "only-ref" in open(tmp_path / "list.txt").read()
    bent = ref[0].split("\t"); bent[5] = O.fmt_g(np.float32(float(bent[5]) + 0.01))
    r = run(["\t".join(bent)] + ref[1:], ref)
    assert r.returncode == 1 and "differences above the tolerance (+ print resolution): 1" in r.stdout


def test_packer_sse2_and_avx2_paths_agree(tmp_path):
    """The host 2-bit packer picks its AVX-512BW or AVX2 (+ BMI2) routine at run time (BLAMM_B200_NO_AVX512=1 / BLAMM_B200_NO_AVX2=1
    step down to AVX2 / SSE2): all must give
    the same code and mask words on every byte value, both lower-case rules, lengths around the 32-character step."""
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from blamm_b200 import capi\n"
        "rng = np.random.default_rng(5)\n"
        "out = []\n"
        "for n in (0, 1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 1000, 4099, 1 << 20):\n"
        "    a = rng.integers(0, 256, size=n, dtype=np.uint8)\n"
        "    b = np.frombuffer(b'ACGTacgtNn', dtype=np.uint8)[rng.integers(0, 10, size=n)]\n"
        "    for block in (a, b):\n"
        "        for lower in (capi.LOWER_ZERO, capi.LOWER_FOLD):\n"
        "            c, z, h = capi.pack_ascii(block, lower)\n"
        "            out += [c, z, np.array([h], dtype=np.uint32)]\n"
        "np.save(sys.argv[1], np.concatenate([o.astype(np.uint32) for o in out]))\n" % ROOT)
    res = []
    for k, env in enumerate(({}, {"BLAMM_B200_NO_AVX512": "1"}, {"BLAMM_B200_NO_AVX2": "1"})):
        f = str(tmp_path / ("p%d.npy" % k))
        subprocess.run([sys.executable, "-c", code, f], check=True, env=dict(os.environ, **env))
        res.append(np.load(f))
    assert len(res[0]) > 100000 and np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])


def test_writer_reports_a_full_file_system(tmp_path):
    """The occurrence file in a tmpfs that runs out of space: the mapped writer ends the run with a message and exit code 1 (a store
    into such a mapping raises SIGBUS), the pwrite writer reports the failed write.  Needs the right to mount a 2 MB tmpfs."""
    mnt = tmp_path / "tiny"
    mnt.mkdir()
    if subprocess.run(["mount", "-t", "tmpfs", "-o", "size=2m", "tmpfs", str(mnt)], capture_output=True).returncode != 0:
        pytest.skip("cannot mount a tmpfs here")
    try:
        r = subprocess.run([CLI, "selftest-writer", "200000", "4"], capture_output=True, text=True, env=dict(os.environ, TMPDIR=str(mnt)), timeout=120)
        assert r.returncode == 1 and "Cannot write to the occurrence file" in r.stderr
        r = subprocess.run([CLI, "selftest-writer", "200000", "4"], capture_output=True, text=True,
                           env=dict(os.environ, TMPDIR=str(mnt), BLAMM_B200_WRITER="pwrite"), timeout=120)
        assert r.returncode == 1 and "identical" not in r.stdout
    finally:
        subprocess.run(["umount", str(mnt)], capture_output=True)
