#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/blamm_dropin: the reference (biointec/blamm) with the three edits INTEGRATION.md asks
a maintainer to make -- two accessors on SeqBlock, the declaration of PWMScan::scanPWMB200, its call where `-c` used to reach
scanPWMCUBLAS -- and integration/scanPWMB200.inc appended, linked against libb200scan.so.  The edits are applied to a TEMPORARY copy
of the reference's sources (removed again; no reference source enters the repository); only the binary stays, in the git-ignored
oracle/_ref/.  usage: build_dropin.py <reference dir> <out dir> <openblas dir> <openblas lib name>"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def insert(lines, index, text):
    lines[index:index] = [text + "\n"]


def main():
    ref, out, ob, oblib = sys.argv[1:5]
    lib = os.path.join(ROOT, "blamm_b200", "lib")
    if not os.path.exists(os.path.join(lib, "libb200scan.so")):
        print("build_dropin: blamm_b200/lib/libb200scan.so absent (run `make` first) - skipped")
        return 0
    tmp = tempfile.mkdtemp(prefix="dropin_src_", dir=out)
    try:
        for f in glob.glob(os.path.join(ref, "src", "*.cpp")) + glob.glob(os.path.join(ref, "src", "*.h")):
            shutil.copy(f, tmp)
        # 1. SeqBlock: two read-only accessors
        p = os.path.join(tmp, "sequence.h")
        L = open(p).readlines()
        c = next(i for i, l in enumerate(L) if l.startswith("class SeqBlock"))
        pub = next(i for i in range(c, len(L)) if L[i].strip() == "public:")
        insert(L, pub + 1, "        const std::string& str() const { return block; }\n"
                           "        const std::map<size_t, SeqPos>& markers() const { return block2seq; }")
        open(p, "w").writelines(L)
        # 2. PWMScan: the declaration, next to the other back ends
        p = os.path.join(tmp, "pwmscan.h")
        L = open(p).readlines()
        c = next(i for i, l in enumerate(L) if l.startswith("class PWMScan"))
        at = next(i for i in range(c, len(L)) if L[i].strip() == "#ifdef HAVE_CUDA")
        insert(L, at, "        void scanPWMB200(size_t speciesID, FastaBatch& seqBatch);")
        open(p, "w").writelines(L)
        # 3. pwmscan.cpp: `-c` reaches scanPWMB200; the "CUDA support not enabled" message of a CPU-only build goes; the stub is appended
        p = os.path.join(tmp, "pwmscan.cpp")
        L = open(p).readlines()
        at = next(i for i, l in enumerate(L) if "} else if (cudaMode) {" in l)
        insert(L, at + 1, "                        scanPWMB200(speciesID++, seqBatch);")
        L = [l for l in L if "CUDA support not enabled" not in l and "recompile with CUDA support" not in l]
        L.append("\n" + open(os.path.join(ROOT, "integration", "scanPWMB200.inc")).read())
        open(p, "w").writelines(L)
        cmd = ["/usr/bin/g++", "-O3", "-std=c++11", "-DNDEBUG", "-DHAVE_CONFIG_H", "-DBLAMM_MAJOR_VERSION=1", "-DBLAMM_MINOR_VERSION=0",
               "-DBLAMM_PATCH_LEVEL=0", "-I" + os.path.join(out, "cfg"), "-I" + os.path.join(ROOT, "include"), "-include", "array", "-w"]
        cmd += sorted(glob.glob(os.path.join(tmp, "*.cpp"))) + ["-o", os.path.join(out, "blamm_dropin")]
        cmd += ["-L" + ob, "-l:" + oblib, "-L" + lib, "-lb200scan", "-Wl,--disable-new-dtags", "-Wl,-rpath," + ob + ":" + lib, "-lpthread"]
        subprocess.check_call(cmd)
        print("build_dropin: built", os.path.join(out, "blamm_dropin"))
        return 0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
