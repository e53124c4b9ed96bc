import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def emu_lib():
    """Sequential emulation build of the engine (kernel bodies as host loops) — test infrastructure."""
    from nlzm_b200 import _lib
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["make", "-s", "-C", d])
    return _lib.bind_prototypes(C.CDLL(os.path.join(d, "libnlzm_mf_emu.so")))


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; GPU tests must go through it (and fail loudly if it is missing)."""
    from nlzm_b200 import _lib
    return _lib.load()


def csr_from_find(off, steps):
    return off.astype(np.uint64), steps["dist"].copy(), steps["len"].copy()
