"""Kernel-logic tests on CPU: the CUDA sources of suffix_array_b200/csrc compiled against the SIMT
emulator of tests/emu (fibers; interleaved blocks, so the decoupled look-back takes both its
PARTIAL and INCLUSIVE paths) and compared bit-exactly with the oracle.  These are NOT the parity
gate -- that is tests/test_gpu_*.py on a B200 -- they keep the kernels honest where no GPU exists."""
import os

import numpy as np
import pytest

from tests import parity_cases as pc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_emu_golden_and_doctests(emu_backend, oracle, golden):
    pc.check_golden(oracle, golden)
    pc.check_doctests(golden)


def test_emu_adversarial(emu_backend, oracle):
    for s in pc.adversarial_texts():
        pc.check_construction(oracle, s)


def test_emu_random_construction(emu_backend, oracle):
    rng = np.random.default_rng(11)
    for trial in range(12):
        n = int(rng.integers(0, 30000))
        sigma = int(rng.choice([1, 2, 4, 5, 16, 100, 256]))
        s = rng.integers(0, sigma, n, dtype=np.uint8)
        if trial % 3 == 0 and n > 100:
            p = int(rng.integers(1, 300))
            s = np.tile(s[:p], n // p + 1)[:n].copy()
            s[int(rng.integers(0, n))] ^= 1
        pc.check_construction(oracle, s)


def test_emu_many_tiles(emu_backend, oracle):
    # > 30 onesweep tiles and > 60 scan tiles: chained look-back across many tiles
    from suffix_array_b200 import gen
    pc.check_construction(oracle, gen.dna_like(150000))
    pc.check_construction(oracle, gen.repetitive(60000, block=700, mut_rate=1e-3))


def test_emu_queries(emu_backend, oracle):
    rng = np.random.default_rng(5)
    for sigma, n in ((4, 5000), (256, 3000), (2, 800), (1, 300)):
        s = rng.integers(0, sigma, n, dtype=np.uint8)
        pats = pc.random_patterns(rng, s, 120, max_len=90) + [b"", s[:1].tobytes(), s[-1:].tobytes(), s.tobytes()]
        pc.check_queries(oracle, s, pats)
    pc.check_queries(oracle, np.frombuffer(b"", dtype=np.uint8), [b"", b"a", b"ab"])


def test_emu_prefix_directory_queries(emu_backend, oracle):
    """The prefix directory must not change a single answer: search_all / contains / search_lcp against the oracle
    on alphabets with gaps and patterns with bytes the text does not contain (see pc.directory_query_cases)."""
    import ctypes as C
    rng = np.random.default_rng(404)
    for s, pats in pc.directory_query_cases(rng):
        sa = pc.check_queries(oracle, s, pats)
        sigma, depth = C.c_uint32(), C.c_uint32()
        entries = emu_backend.sab200_index_directory(sa._get_index(), C.byref(sigma), C.byref(depth))
        assert entries == int(sigma.value) ** int(depth.value) and entries >= 2, (entries, sigma.value, depth.value)


def test_emu_probe_counter_and_directory_switch(emu_backend, oracle, monkeypatch):
    """sab200_index_probes counts the suffix comparisons of search_all; with the prefix directory a batch needs
    fewer of them than with SAB_SEARCH_DIR=0 (bisection of the whole bucket), and both give the oracle's answers."""
    from suffix_array_b200 import SuffixArray
    rng = np.random.default_rng(5)
    s = rng.integers(0, 4, 60000, dtype=np.uint8)
    pats = pc.random_patterns(rng, s, 300, max_len=40)
    flat = np.frombuffer(b"".join(pats), dtype=np.uint8)
    offs = np.zeros(len(pats) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(p) for p in pats])
    exp_sa = oracle.saca(s)
    elo, ehi = oracle.search_all_batch(s, exp_sa, None, flat, offs)
    counts = {}
    for switch in ("1", "0"):
        monkeypatch.setenv("SAB_SEARCH_DIR", switch)
        sa = SuffixArray(s)
        ix = sa._get_index()
        assert (emu_backend.sab200_index_directory(ix, None, None) > 0) == (switch == "1")
        assert emu_backend.sab200_index_probes(ix, 1) == 0
        lo, hi = sa.search_all_batch(flat, offs)
        counts[switch] = emu_backend.sab200_index_probes(ix, 0)
        assert np.array_equal(lo, elo) and np.array_equal(hi, ehi)
    assert 0 < counts["1"] < counts["0"] / 2, counts


def test_emu_fused_buckets(emu_backend, oracle):
    """Bucket table from the sorted keys of the construction: absent bytes, \\0 / \\xff, one symbol per key (falls
    back to the pair counting over the resident text), all 256 byte values, 255 values (radix 2^8: base^k = 2^64)."""
    rng = np.random.default_rng(31)
    texts = [b"", b"a", b"ab", b"banana", b"\x00", b"\xff\x00\xff", b"mississippi" * 5, bytes(range(256)) * 3,
             bytes(range(255)) * 40, bytes(range(1, 256)) * 40, rng.integers(0, 256, 30000, dtype=np.uint8),
             rng.integers(0, 255, 40000, dtype=np.uint8), rng.integers(3, 7, 5000, dtype=np.uint8)]
    for s in texts:
        pc.check_fused_buckets(oracle, s)
    for _ in range(3):
        pc.check_fused_buckets(oracle, pc.random_text(rng))


def test_emu_lcp_array(emu_backend, oracle):
    rng = np.random.default_rng(12)
    for s in (b"", b"a", b"banana", b"mississippi" * 3, b"aaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaaa", b"\x00\xff" * 40):
        pc.check_lcp(oracle, s)
    for _ in range(3):
        pc.check_lcp(oracle, pc.random_text(rng))


def test_emu_from_parts(emu_backend, oracle):
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 50, 3000):
        pc.check_from_parts(oracle, rng.integers(0, 4, n, dtype=np.uint8))


def test_emu_pack(emu_backend, oracle):
    rng = np.random.default_rng(9)
    for n in (0, 1, 5, 126, 127, 128, 129, 255, 256, 1000, 5000):
        pc.check_pack(oracle, rng.integers(0, 256, n, dtype=np.uint8))
    pc.check_pack(oracle, np.frombuffer(b"banana", dtype=np.uint8))
    assert SuffixArrayBanana().hex() == "53413478070000000d0000000000000006000000250000001300000001"


def SuffixArrayBanana():
    from suffix_array_b200 import SuffixArray
    return SuffixArray(b"banana").dump_bytes()


def test_emu_group_sort_boundaries(emu_backend, oracle):
    """Rounds whose groups sit right at the limits of group_sort_kernel: a random block repeated r times gives
    groups of r records (r = 32: counting rank, r = 33: warp sort, r = 600: radix path), lists longer than one
    2048-record tile (groups cut by tile borders), and mixtures of the kinds."""
    from suffix_array_b200 import _lib, gen
    rng = np.random.default_rng(77)
    for r, blk in ((31, 97), (32, 90), (33, 90), (34, 61), (2, 1500), (3, 1100), (130, 40), (600, 9)):
        block = rng.integers(0, 4, blk, dtype=np.uint8)
        t = np.concatenate([np.tile(block, r), rng.integers(0, 4, 50, dtype=np.uint8)])
        pc.check_construction(oracle, t)
        st = _lib.last_stats()
        assert st["group_sort_records"] > 0, (r, blk, st)
    # split filter + in-group sort: parked groups make the next list two ascending runs; the large-group records
    # must then go through the full radix sort (regression: they were returned to positions of other groups)
    pc.check_construction(oracle, gen.repetitive(9518, block=2180, mut_rate=0.01))
    # small and large groups interleaved in one list
    a = np.tile(rng.integers(0, 4, 7, dtype=np.uint8), 700)
    b = np.tile(rng.integers(0, 4, 400, dtype=np.uint8), 3)
    pc.check_construction(oracle, np.concatenate([a, rng.integers(0, 4, 3000, dtype=np.uint8), b]))
    st = _lib.last_stats()
    assert st["group_sort_records"] > st["group_big_records"] > 0, st


def test_emu_radix_sort_direct(emu_lib):
    """The LSD radix sort on its own (through sab200_sort_pairs_device): stable order of (key, payload) pairs
    against numpy for sizes around the tile (256 threads x 18 items = 4608) and look-back group (8 tiles) limits and for skewed,
    constant and wide digits -- the two-level look-back sees complete, partial and single groups."""
    import ctypes as C
    L = emu_lib
    rng = np.random.default_rng(123)
    cases = []
    T = 4608  # SAB_PASS_THREADS * SAB_PASS_ITEMS (csrc/sab_sort.cuh)
    for n in (1, 31, 4095, 4096, 4097, T - 1, T, T + 1, 8 * T, 8 * T + 1, 9 * T - 1, 70000, 17 * T):
        cases.append((n, 64, rng.integers(0, 2 ** 63, n, dtype=np.int64).astype(np.uint64) * np.uint64(2)
                      + rng.integers(0, 2, n, dtype=np.int64).astype(np.uint64)))
    cases.append((50000, 17, rng.integers(0, 2 ** 17, 50000, dtype=np.int64).astype(np.uint64)))          # 3 passes, top one partial
    cases.append((40000, 40, (rng.integers(0, 3, 40000, dtype=np.int64).astype(np.uint64) << np.uint64(32))
                  | rng.integers(0, 5, 40000, dtype=np.int64).astype(np.uint64)))                          # few distinct digits
    cases.append((33000, 64, np.full(33000, 0x0123456789ABCDEF, dtype=np.uint64)))                        # every pass is a no-op
    cases.append((36000, 48, np.sort(rng.integers(0, 2 ** 48, 36000, dtype=np.int64).astype(np.uint64))[::-1].copy()))  # reversed
    for n, bits, keys in cases:
        vals = rng.permutation(n).astype(np.uint32)
        k0, v0 = keys.copy(), vals.copy()
        k1, v1 = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint32)
        which = L.sab200_sort_pairs_device(k0.ctypes.data, k1.ctypes.data, v0.ctypes.data, v1.ctypes.data, n, bits, 0)
        assert which in (0, 1), which
        ks, vs = (k0, v0) if which == 0 else (k1, v1)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(ks, keys[order]), (n, bits)
        assert np.array_equal(vs, vals[order]), (n, bits)


def test_argument_errors(emu_lib):
    """Error behaviour at the boundary: bad sizes / null pointers return SAB200_ERR_ARGS (-1) with
    a message instead of the reference's panics (src/saca.rs:10-11); nothing is dereferenced."""
    import ctypes as C
    L = emu_lib
    buf = np.zeros(16, dtype=np.uint32)
    txt = np.zeros(16, dtype=np.uint8)
    too_long = 0xFFFFFFFF  # MAX_LENGTH + 1
    assert L.sab200_saca(txt.ctypes.data, too_long, buf.ctypes.data, 1) == -1
    assert b"MAX_LENGTH" in L.sab200_last_error() or L.sab200_last_error()
    assert L.sab200_saca(txt.ctypes.data, 4, None, 1) == -1
    assert L.sab200_saca(None, 4, buf.ctypes.data, 1) == -1
    assert L.sab200_saca(txt.ctypes.data, 4, buf.ctypes.data, 2) == -1      # more GPUs than visible (the emulator has one)
    assert L.sab200_saca(txt.ctypes.data, 4, buf.ctypes.data, 17) == -1
    assert L.sab200_enable_buckets(txt.ctypes.data, too_long, buf.ctypes.data) == -1
    assert L.sab200_check(txt.ctypes.data, too_long, buf.ctypes.data, 5) == -1
    assert not L.sab200_index_create(txt.ctypes.data, 4, None, 5, None, 1)
    n_out = C.c_uint64()
    assert L.sab200_pack(buf.ctypes.data, 4, txt.ctypes.data, 3, C.byref(n_out)) == -1        # output buffer too small
    assert L.sab200_unpack(txt.ctypes.data, 8, buf.ctypes.data, 16, C.byref(n_out)) == -1     # truncated header
    # the empty text is legal everywhere (src/tests.rs strategies include it)
    sa0 = np.full(1, 7, dtype=np.uint32)
    assert L.sab200_saca(None, 0, sa0.ctypes.data, 1) == 0 and sa0[0] == 0
    assert L.sab200_check(None, 0, sa0.ctypes.data, 1) == 1


@pytest.mark.parametrize("build", ["emu", "emu-prod"])
def test_emu_randomized_construction(build, emu_lib, oracle, monkeypatch):
    """Fixed-seed randomized constructions under both emulator builds (`emu-prod` = production cost-model
    constant: few suffixes stay active after the initial sort, so the lazy inverse suffix array and the
    in-group sort carry the rounds as they do on the GPU).  This pair found the split-filter / in-group-sort
    ordering bug fixed in round 1."""
    import ctypes
    import subprocess
    from suffix_array_b200 import _lib
    lib = emu_lib
    if build == "emu-prod":
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "suffix_array_b200", "csrc"), "emu-prod"], stdout=subprocess.DEVNULL)
        lib = _lib._bind(ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libsab200_emu_prod.so")))
    monkeypatch.setattr(_lib, "_lib", lib)
    rng = np.random.default_rng(2026)
    for _ in range(9):
        pc.check_construction(oracle, pc.random_text(rng))
    if build == "emu-prod":
        # split filter -> in-group sweep on a two-run list -> large groups in both runs (see the docstring)
        pc.check_construction(oracle, pc.parked_then_unsorted_text(np.random.default_rng(1)))
        st = _lib.last_stats()
        assert st["group_big_records"] > 0 and st["group_sort_records"] > 2 * st["group_big_records"], st


def test_emu_staged_pageable_copies(emu_lib):
    """The bounce-buffer path for pageable caller memory (csrc/sab_api.cu sab_copy_*), forced on with several copy
    threads: sizes whose byte count does not divide by the thread count (a rounding slip once dropped the last
    entry of a 3 MiB text's suffix array -- found by the GPU suite)."""
    import subprocess
    import sys
    code = r'''
import ctypes, sys
import numpy as np
sys.path.insert(0, %r)
from suffix_array_b200 import _lib, SuffixArray
from oracle import oracle
_lib._lib = _lib._bind(ctypes.CDLL(%r))
rng = np.random.default_rng(41)
for n in (1050624, (1 << 20) + 3):  # (n + 1) * 4 = 6 * 4096 * 171 + 4: the case that lost its tail
    s = rng.integers(0, 256, n, dtype=np.uint8)
    assert np.array_equal(SuffixArray(s).sa, oracle.saca(s)), n
print("ok")
''' % (ROOT, os.path.join(ROOT, "tests", "emu", "libsab200_emu_prod.so"))
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "suffix_array_b200", "csrc"), "emu-prod"], stdout=subprocess.DEVNULL)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, SAB_FORCE_STAGED="1", SAB_COPY_THREADS="6"))
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


def test_emu_group_sort_direct(emu_lib):
    """group_sort_kernel on its own (through sab200_group_sort_device) against numpy: every size class, groups cut
    by tile borders (completed by the tile that owns the head), lists that are not ascending in r1."""
    rng = np.random.default_rng(99)
    seen_big = seen_none = 0
    for sizes, r2v, asc in pc.group_sort_cases(rng):
        keys, vals = pc.grouped_records(rng, sizes, r2v, asc)
        nbig = pc.check_group_sort(emu_lib, keys, vals, asc)
        exp_big = int(sum(s for s in sizes if s > 512))
        assert nbig == exp_big, (sizes[:8], nbig, exp_big)
        seen_big += nbig > 0
        seen_none += nbig == 0
    assert seen_big and seen_none
