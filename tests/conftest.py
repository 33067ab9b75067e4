import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def native_build():
    """Make sure the in-tree native libraries exist (cross-compiles without a GPU)."""
    from blamm_b200 import build, lib_dir
    need = [os.path.join(lib_dir(), f) for f in ("libb200scan.so", "libblammhost.so", "blamm-b200")]
    if not all(os.path.exists(p) for p in need):
        build()
    return lib_dir()


@pytest.fixture(scope="session")
def golden():
    return os.path.join(ROOT, "tests", "golden")
