"""The C++ host mirror (include/sab200_suffix_array.hpp): compiles on CPU, runs its reference-style
checks on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_mirror")


def _compile():
    from suffix_array_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"),
                           "-L" + os.path.join(ROOT, "suffix_array_b200"), "-lsab200", "-o", EXE])


def test_cpp_mirror_compiles():
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_runs(gpu_lib):
    _compile()
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "suffix_array_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([EXE], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr
