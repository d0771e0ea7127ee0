import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/sa_oracle.c): the checker, never the thing under test."""
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "sa_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def emu_lib():
    """libsab200_emu.so: the CUDA sources compiled against the SIMT emulator of tests/emu.  Used by
    the CPU-only kernel-logic tests; it is test infrastructure and is never loaded by the package."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "suffix_array_b200", "csrc"), "emu"],
                          stdout=subprocess.DEVNULL)
    from suffix_array_b200 import _lib
    return _lib._bind(ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libsab200_emu.so")))


@pytest.fixture()
def emu_backend(emu_lib, monkeypatch):
    """Points the host mirror (suffix_array_b200.SuffixArray) at the emulator build for one test."""
    from suffix_array_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", emu_lib)
    return emu_lib


@pytest.fixture(scope="session")
def gpu_lib():
    from suffix_array_b200 import _lib
    return _lib.require_gpu()


def pytest_collection_modifyitems(config, items):
    """The multi-GPU cases run last: with `-x` a failure there must not keep the single-GPU parity tests
    (the first gate) from running."""
    items.sort(key=lambda it: 1 if "test_gpu_multi" in it.nodeid else 0)  # stable: the rest keeps its order
