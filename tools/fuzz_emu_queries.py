"""Developer tool: randomized batched queries (search_all / contains / search_lcp, with and without buckets) on the
SIMT-emulator builds against the oracle: random alphabets with gaps, patterns with bytes the text does not contain.
usage: fuzz_emu_queries.py <emu|emu_prod> <seed> <seconds>.  Round 2, final build: 4647 iterations, no failure."""
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
    sig = int(rng.choice([1, 2, 3, 4, 5, 17, 100, 255, 256]))
    alphabet = np.sort(rng.choice(256, sig, replace=False)).astype(np.uint8)
    n = int(rng.integers(0, 90000)) if rng.random() < 0.8 else int(rng.integers(0, 40))
    s = alphabet[rng.integers(0, sig, n)] if n else np.zeros(0, dtype=np.uint8)
    if n > 600 and rng.random() < 0.5:
        s[n // 2:n // 2 + 250] = s[:250]
    pats = pc.random_patterns(rng, s, 150, max_len=70) + [b"", s.tobytes()[-3:], s.tobytes()[:5]]
    for _ in range(40):
        m = int(rng.integers(1, 30)); i = int(rng.integers(0, max(1, n - m + 1)))
        p = bytearray(s[i:i + m].tobytes()) or bytearray(b"x")
        p[int(rng.integers(0, len(p)))] = int(rng.integers(0, 256))
        pats.append(bytes(p))
    try:
        pc.check_queries(oracle, s, pats)
    except Exception as e:
        np.save(os.makedirs('/tmp/fuzz', exist_ok=True) or '/tmp/fuzz/failq_%s_%d.npy' % (which, seed), s)
        print('FAIL', which, seed, sig, n, repr(e)[:200], flush=True)
print('done', which, seed0, it, 'iterations', flush=True)
