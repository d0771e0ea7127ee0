"""The C-ABI library loads and exports every symbol include/sab200.h declares (no compute calls:
this runs where there is no GPU), and the host mirror refuses to work without a device."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for header in ("sab200.h",):
        with open(os.path.join(ROOT, "include", header)) as f:
            src = f.read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(sab200_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_exported():
    from suffix_array_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 28
    for name in names:
        assert hasattr(L, name), "libsab200.so does not export %s" % name
    L.sab200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.sab200_version()


def test_sass_is_sm100a_only():
    from suffix_array_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    from suffix_array_b200 import _lib, SuffixArray, SabError
    L = _lib.lib()
    if L.sab200_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(SabError):
        SuffixArray(b"banana")
    sa = np.zeros(7, dtype=np.uint32)
    t = np.frombuffer(b"banana", dtype=np.uint8)
    rc = L.sab200_saca(t.ctypes.data, 6, sa.ctypes.data, 1)
    assert rc == -3 and b"no CPU fallback" in L.sab200_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "suffix_array_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".rs")):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                for needle in ("import oracle", "from oracle", "liboracle", "oracle_", "oracle/"):
                    assert needle not in txt, "%s reaches for the oracle (%r)" % (fn, needle)


def test_generators_are_deterministic():
    from suffix_array_b200 import gen
    a = gen.dna_like(100000)
    assert np.array_equal(a, gen.dna_like(100000)) and set(np.unique(a)) == {65, 67, 71, 84}
    assert np.array_equal(gen.uniform_bytes(1000)[:800], gen.uniform_bytes(800))
    r = gen.repetitive(50000, block=1000, mut_rate=0.0)
    assert np.array_equal(r[:1000], r[1000:2000])
    p, o = gen.patterns(a, 100)
    assert o[-1] == p.size and ((o[1:] - o[:-1]) >= 8).all() and ((o[1:] - o[:-1]) <= 64).all()
