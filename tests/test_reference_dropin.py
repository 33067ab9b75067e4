"""The drop-in boundary, exercised from the reference's side: oracle/_ref/blamm_dropin is biointec/blamm itself with the three edits
INTEGRATION.md asks a maintainer to make and integration/scanPWMB200.inc appended (oracle/build_dropin.py applies them to a
temporary copy of the reference's sources); its `scan -c` reaches libb200scan.so where scanPWMCUBLAS used to be called.

  * CPU suite: the binary runs on the test-only stand-in library (tests/mock, preloaded) -- the binding itself: blocks and markers
    handed over as the header says, positions mapped back with the reference's own SeqBlock::getSeqPos, lines written by its own
    writeOccToDisk.  Output must equal the UNMODIFIED reference's.
  * GPU suite: the same binary with the real library -- the reference's command line scanning on the B200."""
import os
import shutil
import subprocess

import pytest

from blamm_b200 import lib_dir
from tests import util

ROOT = util.ROOT
DROPIN = os.path.join(ROOT, "oracle", "_ref", "blamm_dropin")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "blamm")
pytestmark = pytest.mark.skipif(not os.path.exists(DROPIN), reason="oracle/_ref/blamm_dropin has not been built (needs /root/reference)")


def _ref_env():
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    ob = os.path.join(ROOT, "oracle", "_ref", "openblas_dir.txt")
    if os.path.exists(ob):
        env["LD_LIBRARY_PATH"] = open(ob).read().strip() + ":" + env.get("LD_LIBRARY_PATH", "")
    return env


MODES = (("pt_rc", ["-rc", "-pt", "0.0001"]), ("pt_fwd", ["-pt", "0.0001"]), ("rt_rc", ["-rc"]), ("at_rc", ["-rc", "-at", "9.5"]))


def _example(golden, work, env, payloads):
    shutil.copytree(os.path.join(golden, "example"), work)
    for mode_key, flags in MODES:
        for payload in payloads:
            e = dict(env, BLAMM_DROPIN_PAYLOAD=str(payload))
            r = subprocess.run([DROPIN, "scan", "-c", "-o", "out.txt"] + flags + ["motifs.jaspar", "sequences.mf"], cwd=work, env=e,
                               capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout[-800:] + r.stderr[-800:]
            got = sorted(open(os.path.join(work, "out.txt")).read().splitlines(True))
            assert got == open(os.path.join(golden, "example", "occ_%s.txt" % mode_key)).read().splitlines(True), (mode_key, payload)
            assert ("Wrote %d matches" % len(got)) in r.stdout


def test_reference_binary_with_the_binding_on_the_stand_in_library(golden, tmp_path):
    """CPU: the example in its four modes, in one block and in blocks of 30,000 and 1,024 characters; then two groups with N runs,
    lower case and a record shorter than a motif against the unmodified reference's BLAS path (motifs of at most 14 positions)."""
    from tests.test_cli_mock import _make_inputs
    mock = os.path.join(ROOT, "tests", "mock", "_build", "libb200scan.so")
    if not os.path.exists(mock):
        from oracle import oracle as O
        O.lib()
        os.makedirs(os.path.dirname(mock), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(ROOT, "tests", "mock", "mock_b200scan.cpp"), "-o", mock,
                               "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-lpthread"])
    env = dict(_ref_env(), LD_PRELOAD=mock)                    # (the binary names libb200scan.so; the preloaded object of that name answers)
    _example(golden, str(tmp_path / "ex"), env, (1 << 25, 30000, 1024))
    work = str(tmp_path / "syn")
    os.makedirs(work)
    _make_inputs(work, 210, n_groups=2)
    for args in (["dict", "seq.mf"], ["hist", "motifs.jaspar", "seq.mf"], ["scan", "-t", "2", "-rc", "-pt", "0.001", "-o", "ref.txt", "motifs.jaspar", "seq.mf"]):
        assert subprocess.run([REF_BIN] + args, cwd=work, env=_ref_env(), capture_output=True, text=True).returncode == 0
    r = subprocess.run([DROPIN, "scan", "-c", "-rc", "-pt", "0.001", "-o", "out.txt", "motifs.jaspar", "seq.mf"], cwd=work,
                       env=dict(env, BLAMM_DROPIN_PAYLOAD="20000"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-800:]
    want = sorted(open(os.path.join(work, "ref.txt")).read().splitlines(True))
    assert len(want) > 500 and sorted(open(os.path.join(work, "out.txt")).read().splitlines(True)) == want


@pytest.mark.gpu
def test_reference_binary_with_the_binding_on_the_gpu(golden, tmp_path):
    """GPU: the reference's own command line, `scan -c`, scoring through libb200scan.so on the B200: the golden occurrence files."""
    env = _ref_env()
    env["LD_LIBRARY_PATH"] = lib_dir() + ":" + env.get("LD_LIBRARY_PATH", "")
    _example(golden, str(tmp_path / "ex"), env, (1 << 25, 30000))
