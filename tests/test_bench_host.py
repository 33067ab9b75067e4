"""bench.py's host-side logic on the CPU: the two arms describe the same configuration, the stream is dealt to the ranks the way
the CLI deals chunks, the threaded host packer equals the single call, and the reference arm runs without loading any native
library of this repository."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

from blamm_b200 import capi, shard, synth
from tests import util

ROOT = util.ROOT
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workload_descriptions_are_shared_by_both_arms():
    for fn, a in ((bench.workload_c2, (8, 100.0, False, 0.0)), (bench.workload_c3, (8, 3.1)), (bench.workload_c4, (8, 1000.0)), (bench.workload_c5, (8, 3.1))):
        d = fn(*a)
        assert set(d) == {"workload", "l2"} and "configs[" in d["workload"]
        assert fn(*a) == d                                                   # pure function of the arguments: both arms print the same dict


def test_stream_is_dealt_in_chunks_with_halo():
    world, block, halo = 4, 1000, 34
    plan = shard.plan_shards(block * world, world, halo, block)
    assert [s.rank for s in plan] == [0, 1, 2, 3] and all(s.n_payload == block for s in plan)
    assert [s.n_total for s in plan] == [block + halo] * 3 + [block]        # the last chunk has nothing behind it
    a = bench.chunk_chars(1, 5000, (0.25, 0.25, 0.25, 0.25))
    b = bench.chunk_chars(1, 1 << 16, (0.25, 0.25, 0.25, 0.25))[:5000]
    assert np.array_equal(a, b)                                              # a chunk's head can be generated alone (the halo of the chunk before)
    assert not np.array_equal(a, bench.chunk_chars(2, 5000, (0.25, 0.25, 0.25, 0.25)))


def test_threaded_host_packer_equals_single_call():
    seq = synth.random_acgt(3_300_077, 5)
    seq[1000:2000] |= 0x20
    seq[77] = ord("N")
    codes, zm, hz = capi.pack_ascii(seq)
    pk = bench.HostPacker(3)
    try:
        c2 = np.zeros(len(codes) + 4, np.uint32); z2 = np.zeros(len(zm) + 4, np.uint32)
        has_zero = pk.pack(seq.ctypes.data, len(seq), c2.ctypes.data, z2.ctypes.data)
        assert has_zero == hz and np.array_equal(c2[:len(codes)], codes) and np.array_equal(z2[:len(zm)], zm)
        fut = pk.pack_async(seq.ctypes.data, len(seq), c2.ctypes.data, z2.ctypes.data)
        assert fut.result() == hz and pk.last_ms > 0
    finally:
        pk.close()


def test_reference_arm_loads_no_repo_library(tmp_path):
    """`bench.py --impl reference` on a small slice: one JSON line with impl = reference, the same config as the GPU arm, a
    compute-only figure, and no libb200scan / libblammhost in the process (checked from inside through /proc/self/maps)."""
    if not os.path.exists(bench.REF_BIN):
        import pytest
        pytest.skip("oracle/_ref/blamm not built")
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--ref-mbp', '1', '--steps', '1', '--warmup', '0']\n"
            "runpy.run_path(%r, run_name='__main__')\n"
            "maps = open('/proc/self/maps').read()\n"
            "print('NATIVE', 'libb200scan' in maps, 'libblammhost' in maps)\n" % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "reference" and line["value"] > 0
    assert line["config"] == bench.workload_c2(1, 100.0, False, 0.0)
    assert line["cpu_baseline"]["compute_only"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "NATIVE False False" in r.stdout
