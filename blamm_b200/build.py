"""In-tree build of the native libraries (wraps the repository Makefile)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lib_dir() -> str:
    return os.path.join(ROOT, "blamm_b200", "lib")


def build(verbose: bool = False) -> None:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (see Makefile); cross-compiles without a GPU."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-s", "-C", ROOT, "all"], stdout=out)
    for f in ("libb200scan.so", "libblammhost.so", "blamm-b200"):
        if not os.path.exists(os.path.join(lib_dir(), f)):
            raise RuntimeError("build did not produce " + f)
