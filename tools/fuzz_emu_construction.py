"""Developer tool: randomized constructions on the SIMT-emulator builds against the oracle (structured repeats, parked
groups, natural-language-like and periodic texts).  usage: fuzz_emu_construction.py <emu|emu_prod> <seed> <seconds>;
failing inputs are saved under /tmp/fuzz.  Round 2, final build: 1924 iterations, no failure."""
import ctypes, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from suffix_array_b200 import _lib, SuffixArray, gen
from oracle import oracle
from tests import parity_cases as pc
which = sys.argv[1]; seed0 = int(sys.argv[2]); budget = float(sys.argv[3])
_lib._lib = _lib._bind(ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'emu', 'libsab200_%s.so' % which)))
t0 = time.time(); it = 0
while time.time() - t0 < budget:
    seed = seed0 * 100000 + it; it += 1
    rng = np.random.default_rng(seed)
    kind = int(rng.integers(0, 7))
    n = int(rng.integers(2000, 120000))
    if kind == 0:
        blk = int(rng.integers(8, 4000)); s = gen.repetitive(n, seed=seed, block=blk, mut_rate=float(rng.choice([0, 1e-4, 1e-3, 1e-2, 5e-2])))
    elif kind == 1:
        sig = int(rng.choice([2, 3, 4, 20, 256])); blk = rng.integers(0, sig, int(rng.integers(3, 300)), dtype=np.uint8)
        copies = int(rng.integers(2, 1500)); parts = []
        for c in range(copies):
            b = blk.copy()
            if rng.random() < 0.3: b[int(rng.integers(0, b.size))] = int(rng.integers(0, sig))
            parts.append(b); parts.append(rng.integers(0, sig, int(rng.integers(0, 20)), dtype=np.uint8))
        s = np.concatenate(parts)[:200000]
    elif kind == 2:
        s = pc.parked_then_unsorted_text(rng, copies=int(rng.integers(520, 900)), zlen=int(rng.integers(30, 90)), ulen=int(rng.integers(5, 60)),
                                         lq=int(rng.integers(3000, 50000)), la=int(rng.integers(100, 3000)), lr=int(rng.integers(1000, 60000)))
    elif kind == 3:
        s = gen.english_like(n, seed=seed, vocab=int(rng.choice([4, 16, 64, 512])))
    elif kind == 4:
        s = np.concatenate([gen.uniform_bytes(n // 3, seed=seed), gen.english_like(n // 2, seed=seed, vocab=32), np.full(int(rng.integers(1, 3000)), 7, dtype=np.uint8)])
    elif kind == 5:
        pat = rng.integers(0, 3, int(rng.integers(1, 40)), dtype=np.uint8); s = np.tile(pat, n // pat.size + 1)[:n].copy()
        for _ in range(int(rng.integers(0, 6))): s[int(rng.integers(0, n))] ^= 1
    else:
        s = pc.random_text(rng)
    s = np.ascontiguousarray(s, dtype=np.uint8)
    try:
        sa = SuffixArray(s).sa
        ok = np.array_equal(sa, oracle.saca(s))
    except Exception as e:
        ok = False; print('EXC', e)
    if not ok:
        np.save(os.makedirs('/tmp/fuzz', exist_ok=True) or '/tmp/fuzz/fail_%s_%d.npy' % (which, seed), s)
        print('FAIL', which, seed, kind, s.size, flush=True)
print('done', which, seed0, it, 'iterations', flush=True)
