"""Kernel-logic tests on CPU: the CUDA sources of suffix_array_b200/csrc compiled against the SIMT
emulator of tests/emu (fibers; interleaved blocks, so the decoupled look-back takes both its
PARTIAL and INCLUSIVE paths) and compared bit-exactly with the oracle.  These are NOT the parity
gate -- that is tests/test_gpu_*.py on a B200 -- they keep the kernels honest where no GPU exists."""
import numpy as np

from tests import parity_cases as pc


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
    groups of r records (r = 32: ordered in shared memory, r = 33: radix path), lists longer than one
    2048-record tile (groups cut by tile borders), and mixtures of both kinds."""
    from suffix_array_b200 import _lib
    rng = np.random.default_rng(77)
    for r, blk in ((31, 97), (32, 90), (33, 90), (34, 61), (2, 1500), (3, 1100)):
        block = rng.integers(0, 4, blk, dtype=np.uint8)
        t = np.concatenate([np.tile(block, r), rng.integers(0, 4, 50, dtype=np.uint8)])
        pc.check_construction(oracle, t)
        st = _lib.last_stats()
        assert st["group_sort_records"] > 0, (r, blk, st)
    # small and large groups interleaved in one list
    a = np.tile(rng.integers(0, 4, 70, dtype=np.uint8), 40)
    b = np.tile(rng.integers(0, 4, 400, dtype=np.uint8), 3)
    pc.check_construction(oracle, np.concatenate([a, rng.integers(0, 4, 3000, dtype=np.uint8), b]))
    st = _lib.last_stats()
    assert st["group_sort_records"] > st["group_big_records"] > 0, st
