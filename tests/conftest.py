import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    """The C-ABI library is built in-tree (git-ignored).  Build it on demand so a fresh checkout can run the CPU
    tests; on the GPU box the prebuilt .so travels with the snapshot."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "triple_accel_b200", "libtriple_accel_b200.so")
    if not os.path.exists(lib) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "triple_accel_b200", "csrc"), "-j8", "-s"])


_ensure_built()
