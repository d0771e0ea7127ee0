#!/usr/bin/env python3
"""Generates tests/golden/sa_vectors.json.

The reference (hucsmn/suffix_array, a Rust crate) holds NO golden suffix arrays and cannot be
built in this image (no Rust toolchain; its SACA is the un-vendored crate cdivsufsort 2.0).
Its only known answers are the doc-tests in src/lib.rs:19-40, carried verbatim under
"doctests" below.  Every other vector here is computed by the *definition* the reference's own
validator enforces (src/sa.rs:72-84: strictly increasing suffixes under Rust slice Ord, with
sa[0] = n for the empty suffix, src/saca.rs:13) using a naive Python sort -- independent of the
C oracle and of the CUDA engine -- and the bucket table by a literal Python transcription of
src/sa.rs:95-116 / 123-144.  Because that order is unique, these are exactly the values
SuffixArray::new returns.

Run:  python tests/golden/make_golden.py   (rewrites sa_vectors.json in place)
"""
import json
import os


def naive_sa(s: bytes):
    n = len(s)
    return sorted(range(n + 1), key=lambda i: s[i:])


def buckets(s: bytes):
    bkt = [0] * (256 * 257 + 1)
    bkt[0] = 1
    n = len(s)
    if n > 0:
        for i in range(n - 1):
            bkt[s[i] * 257 + (s[i + 1] + 1) + 1] += 1
        bkt[s[n - 1] * 257 + 1] += 1
    acc = 0
    for i in range(len(bkt)):
        acc += bkt[i]
        bkt[i] = acc & 0xFFFFFFFF
    return bkt


def get_bucket(bkt, sa_len, pat: bytes):
    if bkt is None:
        return [0, sa_len]
    if len(pat) > 1:
        idx = pat[0] * 257 + (pat[1] + 1) + 1
        return [bkt[idx - 1], bkt[idx]]
    if len(pat) == 1:
        st = pat[0] * 257
        return [bkt[st], bkt[st + 257]]
    return [0, 1]


def search_all_range(s: bytes, sa, pat: bytes):
    """Global [lo, hi) with sa[lo:hi] == SuffixArray::search_all (src/sa.rs:173-204)."""
    hits = [j for j, i in enumerate(sa) if s[i:i + len(pat)] == pat]
    if not hits:
        # lower bound position of pat among the suffixes
        lo = sum(1 for i in sa if s[i:] < pat)
        return [lo, lo]
    assert hits == list(range(hits[0], hits[-1] + 1))
    return [hits[0], hits[-1] + 1]


def lcp(a: bytes, b: bytes):
    k = 0
    while k < len(a) and k < len(b) and a[k] == b[k]:
        k += 1
    return k


def fib(k):
    a, b = b"b", b"a"
    for _ in range(k):
        a, b = b, b + a
    return b


def thue_morse(k):
    return bytes((bin(i).count("1") & 1) + 97 for i in range(1 << k))


TEXTS = [
    b"", b"a", b"aa", b"ab", b"ba", b"banana", b"mississippi", b"abracadabra",
    b"splendid splendor", bytes([0xFF, 0x00, 0xFF, 0x00, 0xFF]), b"aaaaaaaa",
    b"\x00", b"\x00\x00\x00", b"a\x00", b"\x00a", b"\x00\x01\x00\x01\x00",
    b"abababababababababab", b"a" * 127, b"a" * 128, b"a" * 129, bytes(range(256)),
    bytes(reversed(range(256))), fib(10), thue_morse(8),
    b"ACGTACGTACGTTTGACGTACGTACGAAACGT" * 5,
    b"the quick brown fox jumps over the lazy dog " * 6,
]

PATTERNS = [b"", b"a", b"b", b"n", b"an", b"na", b"ba", b"ab", b"z", b"nb", b"ana", b"splend",
            b"splash", b"\x00", b"\x00\x00", b"\xff", b"\xff\x00", b"ACGT", b"the ", b"issi", b"aaaa"]


def main():
    vectors = []
    for s in TEXTS:
        sa = naive_sa(s)
        bkt = buckets(s)
        nz = {}
        prev = 0
        for i, v in enumerate(bkt):  # sparse form: only slots where the running sum changes
            if v != prev:
                nz[str(i)] = v
                prev = v
        queries = []
        for p in PATTERNS:
            queries.append({
                "pat_hex": p.hex(),
                "bucket": get_bucket(bkt, len(sa), p),
                "range": search_all_range(s, sa, p),
                "contains": any(s[i:i + len(p)] == p for i in range(len(s) + 1)),
                "lcp_len": max(lcp(p, s[i:]) for i in range(len(s) + 1)),
            })
        vectors.append({"text_hex": s.hex(), "sa": sa, "bkt_steps": nz, "queries": queries})
    doc = {
        "doctests": {  # /root/reference/src/lib.rs:19-40
            "text": "splendid splendor",
            "contains": ["splend", True],
            "search_all": ["splend", [0, 9]],
            "search_lcp": ["splash", "spl"],
        },
        "vectors": vectors,
    }
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sa_vectors.json")
    with open(out, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes;", len(vectors), "texts")


if __name__ == "__main__":
    main()
