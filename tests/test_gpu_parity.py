"""Parity gate: libsab200.so on a B200 against the oracle, through the C ABI (host mirror
suffix_array_b200.SuffixArray -> include/sab200.h).  Bit-exact on every input.

Mirrors the reference's own tests (/root/reference/src/tests.rs:12-59: conversion, contains,
search_all, search_lcp, each without and with enable_buckets; texts any::<u8>() of 0..4096 bytes)
and adds the adversarial and BASELINE.json-shaped inputs of SURVEY.md section 4."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

from suffix_array_b200 import SuffixArray, gen
from tests import parity_cases as pc

pytestmark = pytest.mark.gpu


def test_native_library_is_loaded(gpu_lib):
    import suffix_array_b200._lib as L
    assert b"sm_100a" in gpu_lib.sab200_version()
    with open("/proc/self/maps") as f:
        assert "libsab200.so" in f.read()
    assert L.lib().sab200_device_count() >= 1


def test_golden_and_doctests(gpu_lib, oracle, golden):
    pc.check_golden(oracle, golden)
    pc.check_doctests(golden)


def test_adversarial(gpu_lib, oracle):
    for s in pc.adversarial_texts():
        pc.check_construction(oracle, s)
        pc.check_from_parts(oracle, s)


@st.composite
def bytes_with_pat(draw, max_size):
    s = draw(st.binary(min_size=0, max_size=max_size))
    n = len(s)
    m = int(n * draw(st.floats(min_value=0.0, max_value=0.999)))
    kind = draw(st.integers(0, 2))
    if kind == 0:
        i = draw(st.integers(0, n - m))
        return s, s[i:i + m]
    if kind == 1:
        i = draw(st.integers(0, n - m))
        junk = draw(st.binary(min_size=0, max_size=m))
        return s, s[i:i + (m - len(junk))] + junk
    return s, draw(st.binary(min_size=m, max_size=m))


_hyp = dict(deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large,
                                                   HealthCheck.function_scoped_fixture])


@settings(max_examples=120, **_hyp)
@given(st.binary(min_size=0, max_size=4096))
def test_conversion_correctness(gpu_lib, oracle, s):
    # src/tests.rs:14-17: new(s).into_parts() -> from_parts must be Some; plus bit-exact vs oracle
    t, sa = SuffixArray.new(s).into_parts()
    assert SuffixArray.from_parts(t, sa) is not None
    assert oracle.check_integrity(s, sa)
    assert np.array_equal(sa, oracle.saca(s))


@settings(max_examples=120, **_hyp)
@given(bytes_with_pat(4096))
def test_query_correctness(gpu_lib, oracle, sp):
    # src/tests.rs:20-59: contains / search_all / search_lcp vs the naive checkers, without and with buckets
    s, pat = sp
    sa = SuffixArray.new(s)
    nc = oracle.naive_contains(s, pat)
    na = np.sort(oracle.naive_search_all(s, pat))
    nl = oracle.naive_search_lcp(s, pat)
    for with_bkt in (False, True):
        if with_bkt:
            sa.enable_buckets()
        assert sa.contains(pat) == nc
        assert np.array_equal(np.sort(sa.search_all(pat)), na)
        r = sa.search_lcp(pat)
        assert s[r.start:r.stop] == pat[:nl]


def test_random_mid_sizes(gpu_lib, oracle):
    rng = np.random.default_rng(2024)
    for trial in range(40):
        n = int(rng.integers(0, 300000))
        sigma = int(rng.choice([1, 2, 3, 4, 5, 16, 100, 256]))
        s = rng.integers(0, sigma, n, dtype=np.uint8)
        if trial % 3 == 0 and n > 100:
            p = int(rng.integers(1, 5000))
            s = np.tile(s[:p], n // p + 1)[:n].copy()
            s[int(rng.integers(0, n))] ^= 1
        pc.check_construction(oracle, s)


def test_queries_vs_oracle(gpu_lib, oracle):
    rng = np.random.default_rng(99)
    for sigma, n in ((4, 200000), (256, 100000), (2, 5000), (1, 700), (20, 50000)):
        s = rng.integers(0, sigma, n, dtype=np.uint8)
        pats = pc.random_patterns(rng, s, 3000, max_len=200) + [b"", s[:1].tobytes(), s[-1:].tobytes(), s[:5000].tobytes()]
        pc.check_queries(oracle, s, pats)
    pc.check_queries(oracle, np.frombuffer(b"", dtype=np.uint8), [b"", b"a", b"ab"])


def test_prefix_directory_queries(gpu_lib, oracle):
    # the prefix directory of the resident index must not change a single answer (alphabets with gaps, bytes the
    # text does not contain at every position, patterns shorter / longer than the directory depth)
    import ctypes as C
    rng = np.random.default_rng(404)
    for s, pats in pc.directory_query_cases(rng):
        sa = pc.check_queries(oracle, s, pats)
        sigma, depth = C.c_uint32(), C.c_uint32()
        entries = gpu_lib.sab200_index_directory(sa._get_index(), C.byref(sigma), C.byref(depth))
        assert entries == int(sigma.value) ** int(depth.value) and entries >= 2


def test_queries_sharded_over_replicas(gpu_lib, oracle):
    # the index is replicated, the patterns are sharded over the GPUs, no collective
    ngpus = min(4, gpu_lib.sab200_device_count())
    if ngpus < 2:
        pytest.skip("needs at least 2 GPUs")
    rng = np.random.default_rng(7)
    s = gen.dna_like(1 << 20)
    pats = pc.random_patterns(rng, s, 5001, max_len=64)
    pc.check_queries(oracle, s, pats, ngpus=ngpus)


def test_fused_buckets(gpu_lib, oracle):
    # new() + enable_buckets() in one call, table from the construction's sorted keys
    rng = np.random.default_rng(41)
    texts = [b"", b"a", b"banana", b"\xff\x00\xff", bytes(range(256)) * 3, bytes(range(255)) * 4000, rng.integers(0, 256, 3 << 20, dtype=np.uint8),
             rng.integers(0, 255, 1 << 20, dtype=np.uint8), gen.dna_like(16 << 20), gen.mixed(8 << 20), gen.repetitive(4 << 20, block=1 << 12)]
    for s in texts:
        pc.check_fused_buckets(oracle, s)
    for _ in range(8):
        pc.check_fused_buckets(oracle, pc.random_text(rng))
    ngpus = min(4, gpu_lib.sab200_device_count())
    if ngpus >= 2:
        for s in (gen.dna_like(16 << 20), gen.mixed(8 << 20), rng.integers(0, 256, 3 << 20, dtype=np.uint8), b"banana"):
            pc.check_fused_buckets(oracle, s, ngpus=ngpus)


def test_lcp_array(gpu_lib, oracle):
    # LCP-array builder (chunked Kasai on the GPU) against the oracle's Kasai
    rng = np.random.default_rng(21)
    for s in (b"", b"a", b"banana", b"mississippi" * 3, b"a" * 5000, b"\x00\xff" * 700):
        pc.check_lcp(oracle, s)
    for _ in range(8):
        pc.check_lcp(oracle, pc.random_text(rng))
    for s in (gen.dna_like(8 << 20), gen.repetitive(4 << 20, block=1 << 14), gen.mixed(4 << 20), np.full(1 << 16, 65, dtype=np.uint8)):
        pc.check_lcp(oracle, s)


def test_pack_correctness(gpu_lib, oracle):
    # src/tests.rs:63-76 (feature "pack") + byte equality with the oracle's restatement of the format
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 127, 128, 129, 4095, 4096, 100000):
        pc.check_pack(oracle, rng.integers(0, 256, n, dtype=np.uint8))
    pc.check_pack(oracle, gen.dna_like(3 << 20))
    assert SuffixArray(b"banana").dump_bytes().hex() == "53413478070000000d0000000000000006000000250000001300000001"


def test_baseline_shapes_16mib_bit_exact(gpu_lib, oracle):
    # the named shapes at a size the oracle finishes in seconds: bit-exact
    n = 16 << 20
    for s in (gen.uniform_bytes(n), gen.dna_like(n), gen.repetitive(n, block=1 << 16), gen.mixed(n)):
        pc.check_construction(oracle, s)


def test_filter_and_group_sort_rounds(gpu_lib, oracle):
    # Repeats with 1 % / 3 % mutations: most suffixes stay active (eager inverse suffix array), the split
    # filter parks unsplit groups while other large groups split, small groups go through the in-group sort.
    # The CPU regression of this combination is tests/test_emu_kernels.py::test_emu_group_sort_boundaries.
    from suffix_array_b200 import last_stats
    n = 12 << 20
    for block, mut in ((n * 10 // 44, 1e-2), (n // 3 + 17, 3e-2), (1 << 13, 1e-2)):  # groups of <= 5, <= 3, ~1500 records
        s = gen.repetitive(n, block=block, mut_rate=mut)
        sa = SuffixArray(s)
        assert oracle.sufcheck(s, sa.sa), (block, mut, last_stats())
        assert SuffixArray.from_parts(s, sa.sa) is not None
    # parked large groups, then the in-group sweep on the two-run list, then large groups in both runs
    for seed in (1, 2, 3):
        pc.check_construction(oracle, pc.parked_then_unsorted_text(np.random.default_rng(seed)))


def test_group_sort_direct(gpu_lib):
    """group_sort_kernel on its own (sab200_group_sort_device) against numpy, device buffers through torch: every
    size class (counting rank <= 32, warp sort <= 512, radix path beyond), groups cut by tile borders, lists
    that are not ascending; plus a 32 Mi-record list of ~256-record groups (the shape of BASELINE configs[2])."""
    import torch
    rng = np.random.default_rng(99)

    def to_dev(a):
        return torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a.view(np.int32)).cuda()

    def from_dev(t):
        a = t.cpu().numpy()
        return a.view(np.uint64) if a.dtype == np.int64 else a.view(np.uint32)

    for sizes, r2v, asc in pc.group_sort_cases(rng):
        keys, vals = pc.grouped_records(rng, sizes, r2v, asc)
        nbig = pc.check_group_sort(gpu_lib, keys, vals, asc, to_dev, from_dev)
        assert nbig == int(sum(x for x in sizes if x > 512)), (sizes[:8], nbig)
    sizes = rng.integers(200, 300, (32 << 20) // 250)
    keys, vals = pc.grouped_records(rng, sizes, 3, True)
    assert pc.check_group_sort(gpu_lib, keys, vals, True, to_dev, from_dev) == 0
    sizes = rng.integers(1, 1200, (8 << 20) // 600)
    keys, vals = pc.grouped_records(rng, sizes, 1 << 20, False)
    assert pc.check_group_sort(gpu_lib, keys, vals, False, to_dev, from_dev) == int(sizes[sizes > 512].sum())


def test_c1_uniform_64mib(gpu_lib, oracle):
    # BASELINE.json configs[0] at full size: linear-time verifiers on CPU and GPU
    s = gen.uniform_bytes(64 << 20)
    sa = SuffixArray(s)
    assert oracle.sufcheck(s, sa.sa)
    assert SuffixArray.from_parts(s, sa.sa) is not None


def test_c3_repetitive_256mib(gpu_lib, oracle):
    # BASELINE.json configs[2] at full size (long LCPs, ~14 doubling rounds)
    s = gen.repetitive(256 << 20)
    sa = SuffixArray(s)
    assert SuffixArray.from_parts(s, sa.sa) is not None      # GPU linear-time check (a6 semantics)
    assert oracle.sufcheck(s, sa.sa)                         # CPU linear-time verifier on the full array
    assert np.array_equal(sa.sa, oracle.saca(s))             # bit-exact against the oracle's own construction
    pc.sampled_order_check(s, sa.sa)
    # stress variants (correctness only): pure periodic and all-equal
    for t in (gen.repetitive(8 << 20, block=1 << 12, mut_rate=0.0), np.full(4 << 20, ord("a"), dtype=np.uint8)):
        pc.check_construction(oracle, t)


def test_c2_dna_1gib(gpu_lib, oracle):
    # BASELINE.json configs[1] at full size: the oracle's linear-time verifier over the whole array (the suffix
    # array of a text is unique, so "valid" is "bit-exact"), the GPU verifier, and a bit-exact diff of a
    # 256 MiB prefix-text construction against the oracle's own construction
    s = gen.dna_like(1 << 30)
    sa = SuffixArray(s)
    assert SuffixArray.from_parts(s, sa.sa) is not None
    assert oracle.sufcheck(s, sa.sa)
    pc.sampled_order_check(s, sa.sa)
    head = np.ascontiguousarray(s[:256 << 20])
    assert np.array_equal(SuffixArray(head).sa, oracle.saca(head))
    # search on the finished 1 GiB index (C5 shape, reduced pattern count): exact (lo, hi) / bool
    sa.enable_buckets()
    pats, offs = gen.patterns(s, 20000)
    lo, hi = sa.search_all_batch(pats, offs)
    elo, ehi = oracle.search_all_batch(s, sa.sa, sa.bkt, pats, offs)
    assert np.array_equal(lo, elo) and np.array_equal(hi, ehi)
    assert np.array_equal(sa.contains_batch(pats, offs), oracle.contains_batch(s, sa.sa, sa.bkt, pats, offs))
    assert np.array_equal(sa.bkt, oracle.enable_buckets(s))
