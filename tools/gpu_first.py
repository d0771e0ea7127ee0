"""First GPU bring-up script (developer tool): parity of sab200_saca against the oracle on a B200
and a first look at timings.  Run through gpurun; prints one line per case."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle
from suffix_array_b200 import gen
from suffix_array_b200._lib import Stats

L = C.CDLL(os.path.join(ROOT, "suffix_array_b200", "libsab200.so"))
L.sab200_saca.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32]
L.sab200_saca.restype = C.c_int32
L.sab200_last_error.restype = C.c_char_p
L.sab200_get_stats.argtypes = [C.POINTER(Stats)]
L.sab200_set_profiling.argtypes = [C.c_int32]


def saca(t):
    sa = np.empty(t.size + 1, dtype=np.uint32)
    t0 = time.time()
    rc = L.sab200_saca(t.ctypes.data if t.size else None, t.size, sa.ctypes.data, 1)
    dt = time.time() - t0
    assert rc == 0, (rc, L.sab200_last_error())
    return sa, dt


def stats():
    s = Stats()
    L.sab200_get_stats(C.byref(s))
    return s.as_dict()


d = json.load(open(os.path.join(ROOT, "tests/golden/sa_vectors.json")))
for v in d["vectors"]:
    s = np.frombuffer(bytes.fromhex(v["text_hex"]), dtype=np.uint8)
    sa, _ = saca(s)
    assert sa.tolist() == v["sa"], (v["text_hex"][:40], sa.tolist()[:20])
print("goldens ok", flush=True)
rng = np.random.default_rng(1)
for trial in range(60):
    n = int(rng.integers(0, 200000))
    sig = int(rng.choice([1, 2, 3, 4, 5, 16, 100, 256]))
    s = rng.integers(0, sig, n, dtype=np.uint8)
    if trial % 3 == 0 and n > 100:
        p = int(rng.integers(1, 5000))
        s = np.tile(s[:p], n // p + 1)[:n].copy()
        s[int(rng.integers(0, n))] ^= 1
    sa, _ = saca(s)
    assert np.array_equal(sa, oracle.saca(s)), (trial, n, sig)
print("random vs oracle ok", flush=True)

L.sab200_set_profiling(1)
cases = [("C1 uniform 64MiB", lambda: gen.uniform_bytes(64 << 20), True),
         ("C3 repetitive 64MiB", lambda: gen.repetitive(64 << 20), True),
         ("C2 dna 256MiB", lambda: gen.dna_like(256 << 20), True),
         ("C2 dna 1GiB", lambda: gen.dna_like(1 << 30), False),
         ("C3 repetitive 256MiB", lambda: gen.repetitive(256 << 20), False)]
for name, mk, verify in cases:
    t0 = time.time()
    t = mk()
    tg = time.time() - t0
    sa, dt = saca(t)      # first call includes arena growth
    sa, dt = saca(t)
    st = stats()
    ok = None
    if verify:
        t0 = time.time()
        ok = oracle.sufcheck(t, sa)
        tv = time.time() - t0
    print(json.dumps({"case": name, "gen_s": round(tg, 2), "e2e_s": round(dt, 4), "device_ms": round(st["total_ms"], 3),
                      "MBps_device": round(t.size / 1e6 / (st["total_ms"] / 1e3), 1), "verified": ok,
                      "rounds": st["rounds"], "active": st["active"], "passes": st["passes"],
                      "pass_ms": round(st["radix_pass_ms"], 3), "pass_GBps": round(st["radix_pass_bytes"] / 1e9 / max(st["radix_pass_ms"], 1e-9) * 1e3, 1),
                      "hist_ms": round(st["hist_ms"], 3), "pack_ms": round(st["pack_ms"], 3), "rank_ms": round(st["rank_ms"], 3),
                      "gather_ms": round(st["gather_ms"], 3), "h2d_ms": round(st["h2d_ms"], 2), "d2h_ms": round(st["d2h_ms"], 2),
                      "launches": st["kernel_launches"]}), flush=True)
    del sa, t
print("done")
