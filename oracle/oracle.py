"""ctypes binding of the CPU oracle (oracle/sa_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py.  The product package
(suffix_array_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
BKT_LEN = 256 * 257 + 1

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "sa_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.oracle_saca.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.oracle_saca.restype = C.c_int
        for name in ("oracle_check_integrity", "oracle_sufcheck"):
            f = getattr(L, name)
            f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
            f.restype = C.c_int
        L.oracle_lcp_array.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.oracle_lcp_array.restype = C.c_int
        L.oracle_enable_buckets.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.oracle_enable_buckets.restype = None
        for name in ("oracle_get_bucket", "oracle_get_top_bucket"):
            f = getattr(L, name)
            f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, _u64p, _u64p]
            f.restype = None
        L.oracle_contains.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.oracle_contains.restype = C.c_int
        for name in ("oracle_search_all", "oracle_search_lcp"):
            f = getattr(L, name)
            f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, _u64p, _u64p]
            f.restype = None
        for name in ("oracle_search_all_batch", "oracle_search_lcp_batch"):
            f = getattr(L, name)
            f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                          C.c_void_p, C.c_void_p]
            f.restype = None
        L.oracle_contains_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_uint64, C.c_void_p]
        L.oracle_contains_batch.restype = None
        L.oracle_pack_bound.argtypes = [C.c_uint64]
        L.oracle_pack_bound.restype = C.c_uint64
        L.oracle_pack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.oracle_pack.restype = C.c_uint64
        L.oracle_unpack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.oracle_unpack.restype = C.c_uint64
        L.oracle_naive_contains.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.oracle_naive_contains.restype = C.c_int
        L.oracle_naive_search_all.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
        L.oracle_naive_search_all.restype = C.c_uint64
        L.oracle_naive_search_lcp.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.oracle_naive_search_lcp.restype = C.c_uint64
    return _LIB


def _bytes(s):
    a = np.frombuffer(bytes(s), dtype=np.uint8) if not isinstance(s, np.ndarray) else s
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.size == 0:  # keep a valid pointer
        a = np.zeros(1, dtype=np.uint8)[:0]
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def saca(s):
    """src/saca.rs:9-15 -> np.uint32[n+1]."""
    t = _bytes(s)
    sa = np.empty(t.size + 1, dtype=np.uint32)
    rc = lib().oracle_saca(_p(t), t.size, _p(sa))
    if rc != 0:
        raise RuntimeError("oracle_saca failed rc=%d" % rc)
    return sa


def check_integrity(s, sa):
    t = _bytes(s)
    sa = np.ascontiguousarray(sa, dtype=np.uint32)
    return bool(lib().oracle_check_integrity(_p(t), t.size, _p(sa), sa.size))


def sufcheck(s, sa):
    t = _bytes(s)
    sa = np.ascontiguousarray(sa, dtype=np.uint32)
    return lib().oracle_sufcheck(_p(t), t.size, _p(sa), sa.size) == 1


_par = None


def saca_parallel(s, threads=0):
    """All-cores CPU construction (oracle/sa_parallel.cpp, OpenMP): the "all cores" baseline of bench.py.
    Returns (sa, threads used)."""
    global _par
    if _par is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboracle_par.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.dirname(path), "liboracle_par.so"], stdout=subprocess.DEVNULL)
        _par = C.CDLL(path)
        _par.oracle_saca_parallel.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        _par.oracle_saca_parallel.restype = C.c_int
        _par.oracle_parallel_threads.restype = C.c_int
    t = _bytes(s)
    sa = np.empty(t.size + 1, dtype=np.uint32)
    rc = _par.oracle_saca_parallel(_p(t), t.size, _p(sa), int(threads))
    assert rc == 0
    return sa, (int(threads) if threads else int(_par.oracle_parallel_threads()))


def lcp_array(s, sa):
    t = _bytes(s)
    sa = np.ascontiguousarray(sa, dtype=np.uint32)
    out = np.empty(sa.size, dtype=np.uint32)
    assert lib().oracle_lcp_array(_p(t), t.size, _p(sa), _p(out)) == 0
    return out


def enable_buckets(s):
    t = _bytes(s)
    bkt = np.empty(BKT_LEN, dtype=np.uint32)
    lib().oracle_enable_buckets(_p(t), t.size, _p(bkt))
    return bkt


def get_bucket(bkt, sa_len, pat, top=False):
    p = _bytes(pat)
    lo, hi = C.c_uint64(), C.c_uint64()
    f = lib().oracle_get_top_bucket if top else lib().oracle_get_bucket
    f(_p(bkt), sa_len, _p(p), p.size, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def contains(s, sa, bkt, pat):
    t, p = _bytes(s), _bytes(pat)
    return bool(lib().oracle_contains(_p(t), t.size, _p(sa), _p(bkt), _p(p), p.size))


def search_all(s, sa, bkt, pat):
    """-> (lo, hi): the reference's returned slice is sa[lo:hi] (src/sa.rs:203)."""
    t, p = _bytes(s), _bytes(pat)
    lo, hi = C.c_uint64(), C.c_uint64()
    lib().oracle_search_all(_p(t), t.size, _p(sa), _p(bkt), _p(p), p.size, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def search_lcp(s, sa, bkt, pat):
    """-> (start, end) text range (src/sa.rs:207)."""
    t, p = _bytes(s), _bytes(pat)
    lo, hi = C.c_uint64(), C.c_uint64()
    lib().oracle_search_lcp(_p(t), t.size, _p(sa), _p(bkt), _p(p), p.size, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def search_all_batch(s, sa, bkt, pats, offs):
    t = _bytes(s)
    pats = _bytes(pats)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    np_ = offs.size - 1
    lo = np.empty(np_, dtype=np.uint32)
    hi = np.empty(np_, dtype=np.uint32)
    lib().oracle_search_all_batch(_p(t), t.size, _p(sa), _p(bkt), _p(pats), _p(offs), np_, _p(lo), _p(hi))
    return lo, hi


def search_lcp_batch(s, sa, bkt, pats, offs):
    t = _bytes(s)
    pats = _bytes(pats)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    np_ = offs.size - 1
    lo = np.empty(np_, dtype=np.uint32)
    hi = np.empty(np_, dtype=np.uint32)
    lib().oracle_search_lcp_batch(_p(t), t.size, _p(sa), _p(bkt), _p(pats), _p(offs), np_, _p(lo), _p(hi))
    return lo, hi


def contains_batch(s, sa, bkt, pats, offs):
    t = _bytes(s)
    pats = _bytes(pats)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    np_ = offs.size - 1
    out = np.empty(np_, dtype=np.uint8)
    lib().oracle_contains_batch(_p(t), t.size, _p(sa), _p(bkt), _p(pats), _p(offs), np_, _p(out))
    return out.astype(bool)


def naive_contains(s, pat):
    t, p = _bytes(s), _bytes(pat)
    return bool(lib().oracle_naive_contains(_p(t), t.size, _p(p), p.size))


def naive_search_all(s, pat):
    t, p = _bytes(s), _bytes(pat)
    out = np.empty(t.size + 1, dtype=np.uint32)
    cnt = lib().oracle_naive_search_all(_p(t), t.size, _p(p), p.size, _p(out))
    return out[:cnt].copy()


def naive_search_lcp(s, pat):
    t, p = _bytes(s), _bytes(pat)
    return int(lib().oracle_naive_search_lcp(_p(t), t.size, _p(p), p.size))


def pack(sa):
    """src/packed_sa.rs:17-53,99-106 -> bytes"""
    sa = np.ascontiguousarray(sa, dtype=np.uint32)
    out = np.empty(int(lib().oracle_pack_bound(sa.size)), dtype=np.uint8)
    n = lib().oracle_pack(_p(sa), sa.size, _p(out))
    if n == 0:
        raise ValueError("oracle_pack failed")
    return out[:n].tobytes()


def unpack(data):
    """src/packed_sa.rs:55-88,117-124 -> np.uint32[length]; ValueError on malformed input"""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    if buf.size < 16:
        raise ValueError("truncated")
    length = int(np.frombuffer(buf[4:8].tobytes(), dtype="<u4")[0])
    sa = np.empty(max(length, 1), dtype=np.uint32)
    n = lib().oracle_unpack(_p(buf), buf.size, _p(sa), sa.size)
    if n == 2 ** 64 - 1:
        raise ValueError("malformed packed suffix array")
    return sa[:n].copy()
