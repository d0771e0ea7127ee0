"""Parity checks shared by the GPU tests (tests/test_gpu_*.py, through libsab200.so on a B200) and
the CPU-only kernel-logic tests (tests/test_emu_*.py, same CUDA sources under the SIMT emulator).
Every check compares the engine with the oracle (bit-exact) on the same input."""
import numpy as np

from suffix_array_b200 import SuffixArray, gen


def adversarial_texts():
    """all-equal bytes, 0x00 runs, period-2, Fibonacci / Thue-Morse, n around 128."""
    def fib(k):
        a, b = b"b", b"a"
        for _ in range(k):
            a, b = b, b + a
        return b
    tm = bytes((bin(i).count("1") & 1) + 97 for i in range(1 << 10))
    out = [b"", b"\x00", b"\x00\x00\x00", b"a\x00", b"\x00a", b"\xff", b"\xff\xff", b"ab" * 300, b"a" * 127, b"a" * 128,
           b"a" * 129, b"a" * 1000, b"\x00" * 777, fib(14), tm, bytes(range(256)) * 3, b"abc" * 500 + b"abd",
           b"\xff\x00" * 200 + b"\xff"]
    return [np.frombuffer(t, dtype=np.uint8) for t in out]


def check_construction(oracle, s):
    s = np.ascontiguousarray(s, dtype=np.uint8)
    sa = SuffixArray(s)
    got = sa.into_parts()[1]
    exp = oracle.saca(s)
    assert got.dtype == np.uint32 and got.size == s.size + 1
    assert np.array_equal(got, exp), "suffix array differs from the oracle (n=%d)" % s.size
    return got


def check_golden(oracle, golden):
    for v in golden["vectors"]:
        s = np.frombuffer(bytes.fromhex(v["text_hex"]), dtype=np.uint8)
        sa = SuffixArray(s)
        assert sa.sa.tolist() == v["sa"]
        for with_bkt in (False, True):
            if with_bkt:
                sa.enable_buckets()
                steps, prev = {}, 0
                for i, x in enumerate(sa.bkt.tolist()):
                    if x != prev:
                        steps[str(i)] = x
                        prev = x
                assert steps == v["bkt_steps"]
            pats = [bytes.fromhex(q["pat_hex"]) for q in v["queries"]]
            lo, hi = sa.search_all_batch(pats)
            cont = sa.contains_batch(pats)
            st, en = sa.search_lcp_batch(pats)
            for k, q in enumerate(v["queries"]):
                assert [int(lo[k]), int(hi[k])] == q["range"], (v["text_hex"][:32], q["pat_hex"], with_bkt)
                assert bool(cont[k]) == q["contains"]
                assert int(en[k]) - int(st[k]) == q["lcp_len"]
                assert bytes(s[int(st[k]):int(en[k])]) == pats[k][:q["lcp_len"]]


def check_doctests(golden):
    d = golden["doctests"]  # /root/reference/src/lib.rs:19-40
    s = d["text"].encode()
    sa = SuffixArray.new(s)
    assert sa.contains(d["contains"][0].encode()) is True
    assert sa.search_all(d["search_all"][0].encode()).tolist() == d["search_all"][1]
    r = sa.search_lcp(d["search_lcp"][0].encode())
    assert s[r.start:r.stop] == d["search_lcp"][1].encode()


def random_patterns(rng, s, count, max_len=40):
    """the three schemes of src/tests.rs:79-102"""
    n = s.size
    pats = []
    for _ in range(count):
        m = int(rng.integers(0, min(max_len, n) + 1))
        kind = int(rng.integers(0, 3))
        if kind == 0 or n == 0:
            i = int(rng.integers(0, n - m + 1))
            p = s[i:i + m].tobytes()
        elif kind == 1:
            i = int(rng.integers(0, n - m + 1))
            j = int(rng.integers(0, m + 1))
            p = s[i:i + m - j].tobytes() + rng.integers(0, 256, j, dtype=np.uint8).tobytes()
        else:
            p = rng.integers(0, 256, m, dtype=np.uint8).tobytes()
        pats.append(p)
    return pats


def check_queries(oracle, s, pats, ngpus=1):
    """search_all / contains / search_lcp, without and with buckets, against the literal oracle."""
    s = np.ascontiguousarray(s, dtype=np.uint8)
    sa = SuffixArray(s)
    sa.use_gpus(ngpus)
    exp_sa = oracle.saca(s)
    assert np.array_equal(sa.sa, exp_sa)
    flat = np.frombuffer(b"".join(pats), dtype=np.uint8)
    offs = np.zeros(len(pats) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(p) for p in pats])
    for with_bkt in (False, True):
        bkt = None
        if with_bkt:
            sa.enable_buckets()
            bkt = oracle.enable_buckets(s)
            assert np.array_equal(sa.bkt, bkt), "bucket table differs"
        lo, hi = sa.search_all_batch(flat, offs)
        elo, ehi = oracle.search_all_batch(s, exp_sa, bkt, flat, offs)
        assert np.array_equal(lo, elo) and np.array_equal(hi, ehi), "search_all ranges differ (buckets=%s)" % with_bkt
        assert np.array_equal(sa.contains_batch(flat, offs), oracle.contains_batch(s, exp_sa, bkt, flat, offs))
        st, en = sa.search_lcp_batch(flat, offs)
        est, een = oracle.search_lcp_batch(s, exp_sa, bkt, flat, offs)
        assert np.array_equal(st, est) and np.array_equal(en, een), "search_lcp ranges differ (buckets=%s)" % with_bkt
    return sa


def check_from_parts(oracle, s):
    s = np.ascontiguousarray(s, dtype=np.uint8)
    good = oracle.saca(s)
    assert SuffixArray.from_parts(s, good) is not None
    assert SuffixArray.from_parts(s, good[:-1]) is None  # src/sa.rs:73-75
    if s.size >= 2:
        rng = np.random.default_rng(s.size)
        bad = good.copy()
        i = int(rng.integers(1, s.size))
        bad[i], bad[i + 1] = bad[i + 1], bad[i]
        assert SuffixArray.from_parts(s, bad) is None
        dup = good.copy()
        dup[i] = dup[i + 1]
        assert SuffixArray.from_parts(s, dup) is None
        oob = good.copy()
        oob[i] = s.size + 5
        assert SuffixArray.from_parts(s, oob) is None
        rot = np.roll(good, 1)
        assert SuffixArray.from_parts(s, rot) is None


def sampled_order_check(s, sa, samples=200000, seed=0):
    """Size-independent property for texts too big for the O(n) CPU verifier to be quick:
    sa is a permutation (checksum + bitmap on a sample) and sampled neighbours are in strict order."""
    n = s.size
    assert sa.size == n + 1 and int(sa[0]) == n
    assert int(sa.astype(np.uint64).sum()) == n * (n + 1) // 2
    rng = np.random.default_rng(seed)
    js = rng.integers(1, n + 1, min(samples, n))
    tb = s.tobytes() if n <= (1 << 28) else None
    for j in js[:2000] if tb is None else js[:20000]:
        a, b = int(sa[j - 1]), int(sa[j])
        x = (tb[a:a + 4096] if tb is not None else s[a:a + 4096].tobytes())
        y = (tb[b:b + 4096] if tb is not None else s[b:b + 4096].tobytes())
        assert x < y or (x == y and len(x) == 4096), (int(j), a, b)


def check_pack(oracle, s):
    """pack_correctness (src/tests.rs:63-76): load_bytes(dump_bytes()) round-trips, dump == dump_bytes; plus
    byte equality with the oracle's restatement of the format and loading of the oracle's bytes."""
    import io
    s = np.ascontiguousarray(s, dtype=np.uint8)
    sa1 = SuffixArray(s)
    b1 = sa1.dump_bytes()
    f = io.BytesIO()
    sa1.dump(f)
    assert f.getvalue() == b1
    assert b1 == oracle.pack(sa1.sa), "packed bytes differ from the oracle (n=%d)" % s.size
    sa2 = SuffixArray.load_bytes(s, b1)
    assert np.array_equal(sa1.sa, sa2.sa)
    assert np.array_equal(oracle.unpack(b1), sa1.sa)
    if s.size >= 2:
        bad = bytearray(b1)
        bad[0] ^= 0xFF  # magic
        try:
            SuffixArray.unchecked_load_bytes(s, bytes(bad))
            raise AssertionError("corrupt magic accepted")
        except ValueError:
            pass
        other = np.roll(s, 1)
        if not np.array_equal(other, s):
            try:
                SuffixArray.load_bytes(other, b1)
                raise AssertionError("suffix array of another text accepted")
            except ValueError:
                pass


def check_fused_buckets(oracle, s, ngpus=1):
    """sab200_saca_buckets: suffix array AND bucket table of one call against the oracle (src/sa.rs:89-119)."""
    t = bytes(s) if not isinstance(s, np.ndarray) else s.tobytes()
    sa = SuffixArray.new_with_buckets(t, ngpus=ngpus)
    assert np.array_equal(sa.sa, oracle.saca(t)), len(t)
    exp = oracle.enable_buckets(t)
    bad = np.flatnonzero(sa.bkt != exp)
    assert bad.size == 0, (len(t), bad[:5], sa.bkt[bad[:5]], exp[bad[:5]])


def check_lcp(oracle, s):
    """sab200_lcp_array against the oracle's Kasai (and, for short texts, the definition itself)."""
    t = bytes(s) if not isinstance(s, np.ndarray) else s.tobytes()
    sa = SuffixArray(t)
    got = sa.lcp_array()
    assert np.array_equal(got, oracle.lcp_array(t, sa.sa)), len(t)
    if len(t) <= 300:
        for j in range(1, len(t) + 1):
            a, b = t[int(sa.sa[j - 1]):], t[int(sa.sa[j]):]
            k = 0
            while k < min(len(a), len(b)) and a[k] == b[k]:
                k += 1
            assert got[j] == k
        assert got[0] == 0


def random_text(rng):
    """Texts with structure: runs, repeats with mutations, mixtures of unique and repetitive regions."""
    from suffix_array_b200 import gen
    kind = int(rng.integers(0, 6))
    n = int(rng.integers(1, 40000))
    if kind == 0:
        return rng.integers(0, int(rng.integers(1, 6)), n, dtype=np.uint8)
    if kind == 1:
        return gen.dna_like(n)
    if kind == 2:
        return gen.repetitive(n, block=int(rng.integers(8, 3000)), mut_rate=float(rng.choice([0, 1e-3, 1e-2, 1e-1])))
    if kind == 3:
        return gen.mixed_range(max(n, 64), 0, max(n, 64))
    if kind == 4:
        a = rng.integers(0, 256, n // 2 + 1, dtype=np.uint8)
        b = np.tile(rng.integers(97, 101, int(rng.integers(3, 90)), dtype=np.uint8), n // 50 + 2)
        return np.concatenate([a, b, a[:n // 5]])
    base = gen.repetitive(max(n, 100), block=int(rng.integers(50, 1500)), mut_rate=float(rng.choice([1e-3, 1e-2])))
    return np.concatenate([base, gen.dna_like(max(n // 3, 10)), base[:n // 2]])


def grouped_records(rng, sizes, r2_values, ascending=True):
    """A round's records: groups of the given sizes, contiguous, keyed (r1 << 32) | r2 with r1 = the position a
    sorted suffix array would give the group head; r2 drawn from `r2_values` distinct values (many ties)."""
    sizes = np.asarray(sizes, dtype=np.int64)
    heads = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64) * np.uint64(3) + np.uint64(1)
    order = np.arange(sizes.size)
    if not ascending:  # two ascending runs, as a split-filter round leaves them
        cut = sizes.size // 3
        order = np.concatenate([order[cut:], order[:cut]])
    r1 = np.repeat(heads[order], sizes[order])
    m = int(sizes.sum())
    r2 = rng.integers(0, max(1, r2_values), m, dtype=np.int64).astype(np.uint64)
    if r2_values > 4:
        r2 = r2 * np.uint64(977) + np.uint64(5)
    keys = (r1 << np.uint64(32)) | r2
    vals = rng.permutation(m).astype(np.uint32)
    return keys, vals


def check_group_sort(lib, keys, vals, ascending=True, to_device=None, from_device=None):
    """sab200_group_sort_device against numpy: keys ordered inside every group, the (key, payload) pairs of a
    group preserved as a multiset.  to_device / from_device move arrays for the GPU build (the emulator's
    'device' memory is host memory)."""
    import ctypes as C
    m = keys.size
    hold = []

    def dev(a):
        if to_device is None:
            b = a.copy()
            hold.append(b)
            return b, b.ctypes.data
        t = to_device(a)
        hold.append(t)
        return t, t.data_ptr()

    k0, pk0 = dev(keys)
    k1, pk1 = dev(np.zeros(m, dtype=np.uint64))
    v0, pv0 = dev(vals)
    v1, pv1 = dev(np.zeros(m, dtype=np.uint32))
    nbig = C.c_uint64()
    which = lib.sab200_group_sort_device(pk0, pk1, pv0, pv1, m, 64, 1 if ascending else 0, 0, C.byref(nbig))
    assert which == (1 if m else 0), (which, lib.sab200_last_error())
    if m == 0:
        return 0
    ok = k1 if from_device is None else from_device(k1)
    ov = v1 if from_device is None else from_device(v1)
    gid = np.cumsum(np.concatenate([[0], (keys[1:] >> np.uint64(32)) != (keys[:-1] >> np.uint64(32))]))
    exp_order = np.lexsort((vals, keys, gid))
    assert np.array_equal(ok, keys[exp_order]), "keys not ordered inside their groups"
    got_order = np.lexsort((ov, ok, gid))
    assert np.array_equal(ov[got_order], vals[exp_order]), "payloads do not follow their keys"
    return int(nbig.value)


def group_sort_cases(rng):
    """(sizes, distinct second ranks, ascending) around every limit of group_sort_kernel: counting rank <= 32,
    warp sort 64 / 128 / 256 / 512, the radix path beyond; groups cut by the 2048-record tile borders."""
    cases = []
    for s in (1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 512, 513, 700, 2048, 2049, 5000):
        cases.append(([s] * max(3, 6000 // s), 1 << 20, True))
        cases.append(([s] * max(3, 4200 // s) + [1, 2, 3], 3, True))
    cases.append((rng.integers(1, 40, 900).tolist(), 50, True))
    cases.append((rng.integers(1, 600, 60).tolist(), 7, True))
    cases.append((rng.integers(200, 300, 70).tolist(), 1 << 16, True))
    cases.append((rng.integers(1, 1100, 40).tolist(), 1 << 16, True))
    cases.append(([2047, 1, 2048, 1, 511, 1537, 512, 1536, 513, 1535, 2047 + 512, 1, 2046, 514], 5, True))
    cases.append(([1536, 1024, 1024, 512, 3000, 32, 33, 2015, 33], 1 << 10, True))
    cases.append(([256] * 100, 2, True))
    cases.append(([1], 1, True))
    cases.append(([5], 1, True))
    cases.append(([], 1, True))
    for c in list(cases[-9:-3]) + [cases[20], cases[30]]:
        cases.append((c[0], c[1], False))
    return cases


def directory_query_cases(rng):
    """(text, patterns) aimed at the prefix directory of the resident index (csrc/sab_search.cuh PrefixDir): alphabets
    with gaps, patterns shorter / longer than the directory depth, patterns that end the text, bytes that do not
    occur in the text below / between / above the symbols at every position, the empty pattern."""
    out = []
    for alphabet, n in (([5, 9, 200], 5000), ([0, 1], 3000), ([255], 400), ([0], 300), ([0, 255], 2500),
                        ([7, 8, 9, 10, 250], 70000), (list(range(256)), 20000), ([65, 67, 71, 84], 200000)):
        alphabet = np.array(alphabet, dtype=np.uint8)
        s = alphabet[rng.integers(0, alphabet.size, n)]
        if n > 1000:
            s[n // 2:n // 2 + 300] = s[:300]  # a repeat, so that ranges are longer than one suffix
        pats = [b"", s.tobytes()[-1:], s.tobytes()[-5:], s.tobytes()[-40:], s.tobytes()[:70]]
        absent = [b for b in (0, 1, 4, 6, 8, 66, 100, 199, 201, 254, 255) if b not in alphabet.tolist()]
        for _ in range(150):
            m = int(rng.integers(1, 48))
            i = int(rng.integers(0, n - m + 1))
            p = bytearray(s[i:i + m].tobytes())
            kind = int(rng.integers(0, 4))
            if kind == 1 and absent:
                p[int(rng.integers(0, m))] = int(rng.choice(absent))
            elif kind == 2:
                p[int(rng.integers(0, m))] = int(rng.choice(alphabet))
            elif kind == 3 and absent:
                p = p[:int(rng.integers(0, m))] + bytes([int(rng.choice(absent))])
            pats.append(bytes(p))
        for b in absent[:6]:
            pats.append(bytes([b]))
            pats.append(bytes([b, int(alphabet[0])]))
            pats.append(bytes([int(alphabet[-1]), b]))
        out.append((s, pats))
    return out


def parked_then_unsorted_text(rng, copies=600, zlen=64, ulen=40, lq=60000, la=1200, lr=100000):
    """A text whose doubling rounds go: split filter parks large groups (600 copies of a block whose copies only
    differ late) -> the next round sorts the two-run list with the in-group sweep -> a later round still holds
    large groups in BOTH runs (the parked copies, high first ranks, in front of a run of 1200 equal bytes, low first
    rank).  Regression for the round-2 bug where the list was taken for ascending again after a sweep round and the
    radix-sorted records of large groups went back to the positions of other groups.  ~224 kB: seconds under the
    emulator (production cost-model constant), milliseconds on the GPU."""
    A, B, Cc, D = 97, 98, 99, 100
    Z = rng.choice([Cc, D], zlen).astype(np.uint8)
    P = np.concatenate([np.concatenate([Z, rng.choice([Cc, D], ulen).astype(np.uint8)]) for _ in range(copies)])
    runs = []
    for _ in range(lq // 300):  # periodic runs: groups of ~50 that lose a few members every round (never parked)
        pat = rng.choice([A, B], int(rng.integers(5, 9))).astype(np.uint8)
        runs.append(np.tile(pat, 300 // pat.size + 1)[:300])
        runs.append(rng.choice([Cc, D], 3).astype(np.uint8))
    Q = np.concatenate(runs)
    S = np.full(la, A, dtype=np.uint8)
    R = rng.choice([A, B, Cc, D], lr).astype(np.uint8)  # settled by the initial sort: spare room for the sweep
    return np.concatenate([R, np.array([D], dtype=np.uint8), S, np.array([B], dtype=np.uint8), Q, P])
