"""Synthetic workloads of BASELINE.json.  Every generator is a pure function of
(seed, n) built on splitmix64, so a C/C++ harness can reproduce the same bytes.

The reference's own bench corpus (benches/utils.rs:17-45) is random bytes (alphabet 0..=255) and
Pizza&Chili dna/english downloads; there is no network here, so the DNA- and English-like texts
are synthetic imitations of those shapes.
"""
import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)

SEED_C1 = 0xB2000001
SEED_C2 = 0xB2000002
SEED_C3 = 0xB2000003
SEED_C4 = 0xB2000004
SEED_C5 = 0xB2000005


def splitmix64(seed: int, start: int, count: int) -> np.ndarray:
    """Outputs start .. start+count-1 of the splitmix64 stream seeded with `seed` (uint64[count])."""
    with np.errstate(over="ignore"):
        i = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + i * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def _stream_bytes(seed: int, n: int, out: np.ndarray = None, chunk_words: int = 1 << 22) -> np.ndarray:
    """n bytes = little-endian bytes of successive splitmix64 outputs."""
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    words = (n + 7) // 8
    pos = 0
    w = 0
    while w < words:
        c = min(chunk_words, words - w)
        b = splitmix64(seed, w, c).view(np.uint8)
        take = min(b.size, n - pos)
        out[pos:pos + take] = b[:take]
        pos += take
        w += c
    return out


def uniform_bytes(n: int, seed: int = SEED_C1) -> np.ndarray:
    """C1: uniform random bytes, sigma = 256 (the reference's `random-*` samples, benches/utils.rs:17-22)."""
    return _stream_bytes(seed, n)


_DNA_LUT = np.empty(256, dtype=np.uint8)
_DNA_LUT[:74] = ord("A")
_DNA_LUT[74:128] = ord("C")
_DNA_LUT[128:182] = ord("G")
_DNA_LUT[182:] = ord("T")


def dna_like(n: int, seed: int = SEED_C2, plant: bool = True) -> np.ndarray:
    """C2: A/C/G/T with p = (0.29, 0.21, 0.21, 0.29) plus planted approximate repeats: n/65536
    segments of length 64..4096 copied src -> dst with 1 % point mutations (imitates the LCP tail
    of Pizza&Chili dna, which benches/utils.rs:26-33 would download)."""
    t = _stream_bytes(seed, n)
    np.take(_DNA_LUT, t, out=t)
    if plant and n >= 8192:
        reps = n // 65536
        r = splitmix64(seed ^ 0x5EED5EED, 0, 3 * reps + 3)
        acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
        for k in range(reps):
            ln = 64 + int(r[3 * k] % np.uint64(4033))
            src = int(r[3 * k + 1] % np.uint64(n - ln))
            dst = int(r[3 * k + 2] % np.uint64(n - ln))
            seg = t[src:src + ln].copy()
            mr = splitmix64(seed + 7919 * (k + 1), 0, ln)
            mut = (mr % np.uint64(100)) == 0
            seg[mut] = acgt[((mr[mut] >> np.uint64(8)) % np.uint64(4)).astype(np.int64)]
            t[dst:dst + ln] = seg
    return t


def repetitive(n: int, seed: int = SEED_C3, block: int = 1 << 20, mut_rate: float = 1e-4) -> np.ndarray:
    """C3: a uniform-random block repeated to length n, then n*mut_rate point mutations
    (long LCPs: stresses the doubling rounds and the compaction of settled groups)."""
    block = max(1, min(block, n)) if n else 1
    base = _stream_bytes(seed, block)
    t = np.tile(base, n // block + 1)[:n].copy()
    muts = int(n * mut_rate)
    if muts:
        r = splitmix64(seed ^ 0xA5A5A5A5, 0, 2 * muts)
        pos = (r[:muts] % np.uint64(n)).astype(np.int64)
        t[pos] = (r[muts:] & np.uint64(0xFF)).astype(np.uint8)
    return t


_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_W = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0,
                      1.9, 1.5, 1.0, 0.8, 0.15, 0.15, 0.1, 0.07])


_ENG_BLOCK = 1 << 20
_ENG_CACHE = {}


def _english_tables(seed: int, vocab: int):
    key = (seed, vocab)
    if key not in _ENG_CACHE:
        r = splitmix64(seed ^ 0x0E0E0E0E, 0, vocab * 11)
        lens = (2 + (r[:vocab] % np.uint64(9))).astype(np.int64)
        cum = np.cumsum(_LETTER_W / _LETTER_W.sum())
        u = (r[vocab:vocab * 11] >> np.uint64(11)).astype(np.float64) / float(1 << 53)
        letters = _LETTERS[np.minimum(np.searchsorted(cum, u), 25)].reshape(vocab, 10)
        padded = np.full((vocab, 11), ord(" "), dtype=np.uint8)
        padded[:, :10] = letters
        valid = np.zeros((vocab, 11), dtype=bool)
        for L in range(2, 11):
            rows = lens == L
            valid[rows, :L] = True
            valid[rows, L] = True  # the separating space
            padded[rows, L] = ord(" ")
        zipf = 1.0 / np.arange(1, vocab + 1)
        zcum = np.cumsum(zipf / zipf.sum())
        _ENG_CACHE[key] = (padded, valid, zcum)
    return _ENG_CACHE[key]


def _english_block(seed: int, vocab: int, j: int) -> np.ndarray:
    """Block j (1 MiB) of the English-like text: a pure function of (seed, j)."""
    padded, valid, zcum = _english_tables(seed, vocab)
    out = np.empty(_ENG_BLOCK, dtype=np.uint8)
    pos = 0
    w = 0
    chunk = 1 << 18
    bseed = (seed * 0x9E3779B1 + 0x632BE5AB * (j + 1)) & 0xFFFFFFFFFFFFFFFF
    while pos < _ENG_BLOCK:
        rr = splitmix64(bseed, w, chunk)
        w += chunk
        uu = (rr >> np.uint64(11)).astype(np.float64) / float(1 << 53)
        sel = np.minimum(np.searchsorted(zcum, uu), vocab - 1)
        piece = padded[sel].ravel()[valid[sel].ravel()]
        take = min(piece.size, _ENG_BLOCK - pos)
        out[pos:pos + take] = piece[:take]
        pos += take
    return out


def english_like_range(lo: int, hi: int, seed: int = SEED_C4, vocab: int = 4096) -> np.ndarray:
    """Bytes [lo, hi) of the (unbounded) English-like text: words from a synthetic Zipf(1.0) vocabulary
    (lengths 2..10, letters by English frequency) joined by single spaces, generated in independent
    1 MiB blocks so any range can be produced without the rest (imitates Pizza&Chili english)."""
    out = np.empty(max(hi - lo, 0), dtype=np.uint8)
    p = lo
    while p < hi:
        j = p // _ENG_BLOCK
        blk = _english_block(seed, vocab, j)
        a = p - j * _ENG_BLOCK
        take = min(_ENG_BLOCK - a, hi - p)
        out[p - lo:p - lo + take] = blk[a:a + take]
        p += take
    return out


def english_like(n: int, seed: int = SEED_C4, vocab: int = 4096) -> np.ndarray:
    return english_like_range(0, n, seed, vocab)


def uniform_range(lo: int, hi: int, seed: int = SEED_C1) -> np.ndarray:
    """Bytes [lo, hi) of the uniform byte stream of `seed`."""
    w0 = lo // 8
    w1 = (hi + 7) // 8
    out = np.empty(max(hi - lo, 0), dtype=np.uint8)
    pos = lo
    w = w0
    while w < w1:
        c = min(1 << 22, w1 - w)
        b = splitmix64(seed, w, c).view(np.uint8)
        a = max(pos - w * 8, 0)
        take = min(b.size - a, hi - pos)
        out[pos - lo:pos - lo + take] = b[a:a + take]
        pos += take
        w += c
    return out


def mixed_range(n: int, lo: int, hi: int, seed: int = SEED_C4) -> np.ndarray:
    """Bytes [lo, hi) of the C4 text of total length n: first half uniform bytes, second half
    English-like text.  Every rank of a multi-GPU run generates just its own shard."""
    h = n // 2
    hi = min(hi, n)
    out = np.empty(max(hi - lo, 0), dtype=np.uint8)
    if lo < h:
        e = min(hi, h)
        out[:e - lo] = uniform_range(lo, e, seed)
    if hi > h:
        s0 = max(lo, h)
        out[s0 - lo:] = english_like_range(s0 - h, hi - h, seed)
    return out


def mixed(n: int, seed: int = SEED_C4) -> np.ndarray:
    """C4: first half uniform bytes, second half English-like text."""
    return mixed_range(n, 0, n, seed)


def patterns(text: np.ndarray, count: int, seed: int = SEED_C5, min_len: int = 8, max_len: int = 64,
             hybrid_fraction: float = 0.5, alphabet: np.ndarray = None):
    """C5: `count` patterns of length uniform in [min_len, max_len]: "select" patterns are substrings
    at uniform offsets (always hit); "hybrid" ones keep the first half and replace the second half by
    random symbols drawn from `alphabet` (default: the byte values present in the text) -- the
    reference's pattern schemes, benches/utils.rs:47-60,176-201.  Returns (bytes uint8[], offs uint64[count+1])."""
    n = int(text.size)
    r = splitmix64(seed, 0, 3 * count)
    span = np.uint64(max_len - min_len + 1)
    lens = (np.uint64(min_len) + r[:count] % span).astype(np.int64)
    lens = np.minimum(lens, max(n, 0))
    starts = (r[count:2 * count] % np.maximum(np.uint64(1), (np.uint64(n) - lens.astype(np.uint64) + np.uint64(1)))).astype(np.int64)
    offs = np.zeros(count + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    total = int(offs[-1])
    rep = np.repeat(np.arange(count, dtype=np.int64), lens)
    within = np.arange(total, dtype=np.int64) - np.repeat(offs[:-1].astype(np.int64), lens)
    pats = text[np.repeat(starts, lens) + within].copy()
    if hybrid_fraction > 0 and total:
        if alphabet is None:
            alphabet = np.flatnonzero(np.bincount(text[: min(n, 1 << 24)], minlength=256)).astype(np.uint8)
        is_h = (r[2 * count:3 * count] >> np.uint64(40)).astype(np.float64) / float(1 << 24) < hybrid_fraction
        junk = is_h[rep] & (within >= (lens[rep] + 1) // 2)
        jr = splitmix64(seed ^ 0x77777777, 0, total)
        pats[junk] = alphabet[(jr[junk] % np.uint64(alphabet.size)).astype(np.int64)]
    return pats, offs
