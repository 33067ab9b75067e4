"""Host-side model (C++ through the ctypes C ABI and the blamm-b200 command line) against the golden
fixtures produced by the compiled reference and against the oracle.  No GPU, no compute calls into
libb200scan.so -- only that it loads and exports what include/b200scan.h declares."""
import ctypes
import os
import re
import shutil
import subprocess

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
    assert L.b200scan_abi_version() == 1


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
