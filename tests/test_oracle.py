"""The oracle against the reference's known answers and its own definition (CPU only).

Pins oracle/sa_oracle.c: the doc-test answers of /root/reference/src/lib.rs:19-40,
the hand-verified vectors of tests/golden (naive Python sort, independent of the C code), the
reference's validator (src/sa.rs:72-84, restated literally) and the five properties of
src/tests.rs:12-77 against the reference's own naive checkers (src/tests.rs:104-132).
"""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st


def test_doctests(oracle, golden):
    d = golden["doctests"]
    s = d["text"].encode()
    sa = oracle.saca(s)
    assert oracle.contains(s, sa, None, d["contains"][0].encode()) is d["contains"][1]
    lo, hi = oracle.search_all(s, sa, None, d["search_all"][0].encode())
    assert sa[lo:hi].tolist() == d["search_all"][1]  # SA order, src/lib.rs:29
    a, b = oracle.search_lcp(s, sa, None, d["search_lcp"][0].encode())
    assert s[a:b] == d["search_lcp"][1].encode()
    bkt = oracle.enable_buckets(s)
    lo, hi = oracle.search_all(s, sa, bkt, b"splend")
    assert sa[lo:hi].tolist() == [0, 9]


def test_golden_vectors(oracle, golden):
    for v in golden["vectors"]:
        s = bytes.fromhex(v["text_hex"])
        sa = oracle.saca(s)
        assert sa.tolist() == v["sa"]
        assert sa[0] == len(s)  # src/saca.rs:13
        assert oracle.check_integrity(s, sa) and oracle.sufcheck(s, sa)
        bkt = oracle.enable_buckets(s)
        steps, prev = {}, 0
        for i, x in enumerate(bkt.tolist()):
            if x != prev:
                steps[str(i)] = x
                prev = x
        assert steps == v["bkt_steps"]
        assert bkt[-1] == len(s) + 1
        for q in v["queries"]:
            p = bytes.fromhex(q["pat_hex"])
            assert list(oracle.get_bucket(bkt, len(sa), p)) == q["bucket"]
            for b in (None, bkt):
                assert list(oracle.search_all(s, sa, b, p)) == q["range"] or (
                    # an empty result may sit anywhere inside an empty bucket
                    q["range"][0] == q["range"][1] and len(set(oracle.search_all(s, sa, b, p))) == 1)
                assert oracle.contains(s, sa, b, p) == q["contains"]
                a, e = oracle.search_lcp(s, sa, b, p)
                assert e - a == q["lcp_len"] and s[a:e] == p[:e - a]


def test_bucket_goldens_banana(oracle):
    # SURVEY.md 8c, literal transcription of src/sa.rs:95-116,123-144
    s = b"banana"
    sa = oracle.saca(s)
    bkt = oracle.enable_buckets(s)
    assert bkt[-1] == 7 and bkt[97 * 257] == 1 and bkt[97 * 257 + 1] == 2
    exp = {b"": (0, 1), b"a": (1, 4), b"b": (4, 5), b"n": (5, 7), b"an": (2, 4), b"na": (5, 7), b"ba": (4, 5),
           b"ab": (2, 2), b"z": (7, 7), b"nb": (7, 7)}
    for p, r in exp.items():
        assert oracle.get_bucket(bkt, len(sa), p) == r


def test_validators_reject(oracle):
    s = b"mississippi"
    sa = oracle.saca(s)
    bad = sa.copy()
    bad[3], bad[4] = bad[4], bad[3]
    assert not oracle.check_integrity(s, bad) and not oracle.sufcheck(s, bad)
    assert not oracle.check_integrity(s, sa[:-1]) and not oracle.sufcheck(s, sa[:-1])
    dup = sa.copy()
    dup[5] = dup[6]
    assert not oracle.check_integrity(s, dup) and not oracle.sufcheck(s, dup)


bytes_strategy = st.binary(min_size=0, max_size=600)


@st.composite
def bytes_with_pat(draw, max_size=600):
    """src/tests.rs:79-102: no_junk / trail_junk / all_junk patterns."""
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


@settings(max_examples=150, deadline=None)
@given(bytes_strategy)
def test_conversion_correctness(oracle, s):
    # src/tests.rs:14-17
    sa = oracle.saca(s)
    assert oracle.check_integrity(s, sa)
    assert oracle.sufcheck(s, sa)
    assert sa.tolist() == sorted(range(len(s) + 1), key=lambda i: s[i:])


@settings(max_examples=150, deadline=None)
@given(bytes_with_pat())
def test_query_properties(oracle, sp):
    # src/tests.rs:20-59, without and with buckets
    s, pat = sp
    sa = oracle.saca(s)
    bkt = oracle.enable_buckets(s)
    nc = oracle.naive_contains(s, pat)
    na = np.sort(oracle.naive_search_all(s, pat))
    nl = oracle.naive_search_lcp(s, pat)
    for b in (None, bkt):
        assert oracle.contains(s, sa, b, pat) == nc
        lo, hi = oracle.search_all(s, sa, b, pat)
        assert np.array_equal(np.sort(sa[lo:hi]), na)
        a, e = oracle.search_lcp(s, sa, b, pat)
        assert s[a:e] == pat[:nl]


def test_small_alphabets_and_periodic(oracle):
    rng = np.random.default_rng(7)
    for sigma in (1, 2, 3, 4):
        for n in (1, 2, 3, 7, 64, 127, 128, 129, 1000):
            s = rng.integers(0, sigma, n, dtype=np.uint8).tobytes()
            sa = oracle.saca(s)
            assert sa.tolist() == sorted(range(n + 1), key=lambda i: s[i:])
    s = (b"abcab" * 400)[:1999]
    assert oracle.check_integrity(s, oracle.saca(s))


def _pack_model(sa):
    """Independent pure-Python model of the `pack` byte layout (bincode fixint little-endian
    header "SA4x" + length u32 + data length u64, then BitPacker4x blocks: value j of a 128-value block lives
    in 32-bit lane j % 4 at bit (j // 4) * bits of that lane's LSB-first bit stream, the four lanes interleaved
    word by word; trailing zero bytes of a partial last block are dropped)."""
    import struct
    n = len(sa)
    bits = 0 if n <= 1 else (n - 1).bit_length()
    data = bytearray()
    for b0 in range(0, n, 128):
        blk = list(sa[b0:b0 + 128]) + [0] * (128 - len(sa[b0:b0 + 128]))
        lanes = [0, 0, 0, 0]
        for j, v in enumerate(blk):
            lanes[j % 4] |= (int(v) & ((1 << bits) - 1)) << ((j // 4) * bits)
        out = bytearray()
        for w in range(bits):
            for lane in range(4):
                out += struct.pack("<I", (lanes[lane] >> (32 * w)) & 0xFFFFFFFF)
        if b0 + 128 > n:  # partial block
            while out and out[-1] == 0:
                out.pop()
        data += out
    return struct.pack("<IIQ", 0x78344153, n, len(data)) + bytes(data)


def test_pack_layout_against_python_model(oracle):
    # worked example of SURVEY.md 5.4 ("banana"), then random arrays around the block size
    assert oracle.pack(oracle.saca(b"banana")).hex() == "53413478070000000d0000000000000006000000250000001300000001"
    rng = np.random.default_rng(11)
    for n in (0, 1, 2, 3, 127, 128, 129, 255, 256, 257, 1000, 4096):
        sa = rng.permutation(n).astype(np.uint32) if n else np.zeros(0, dtype=np.uint32)
        b = oracle.pack(sa)
        assert b == _pack_model(sa), n
        assert np.array_equal(oracle.unpack(b), sa)
    for bad in (b"", b"SA4x", _pack_model([1, 0])[:-1], b"XXXX" + _pack_model([1, 0])[4:]):
        with pytest.raises(ValueError):
            oracle.unpack(bad)


def test_parallel_cpu_baseline_matches(oracle):
    """oracle/sa_parallel.cpp (the all-cores CPU baseline of bench.py) against the single-thread oracle."""
    from suffix_array_b200 import gen
    rng = np.random.default_rng(77)
    texts = [b"", b"a", b"banana", b"\x00\x00\x00", b"a\x00", b"a" * 300, b"mississippi" * 9, gen.dna_like(200000).tobytes(),
             gen.repetitive(100000, block=777, mut_rate=1e-3).tobytes(), rng.integers(0, 256, 50000, dtype=np.uint8).tobytes()]
    for t in texts:
        sa, threads = oracle.saca_parallel(t)
        assert threads >= 1 and np.array_equal(sa, oracle.saca(t)), len(t)
