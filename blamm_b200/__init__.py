"""blamm_b200 -- B200-native implementation of the `blamm scan` hot path.

The product is native: blamm_b200/csrc (sm_100a CUDA kernels + the C ABI of include/b200scan.h) and
blamm_b200/host (C++ host model + the blamm-b200 command line).  This Python package only holds the ctypes
bindings used by tests and bench.py, the seeded synthetic-input generators and the chunk-sharding helper.
"""
from .build import build, lib_dir  # noqa: F401
