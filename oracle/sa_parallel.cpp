// sa_parallel.cpp -- TEST / BASELINE INFRASTRUCTURE, never part of the product: an all-cores CPU suffix-array
// construction for the "all cores" column of the benchmark (north_star names psacak, the author's parallel
// SACA-K crate mentioned at /root/reference/README.md:10; it is not a dependency of the reference, not in the
// tree and not installable here, so this file is a from-scratch OpenMP port of prefix doubling -- the same
// algorithm family as the GPU engine -- and is labelled "port" wherever its numbers appear).
//
//   key[i]  = the first 7 symbols of suffix i at 9 bits each (byte + 1; 0 = past the end, so a proper prefix
//             sorts first and real 0x00 bytes stay distinct from padding: /root/reference/src/sa.rs:77-79)
//   sort (key, i) with the parallel-mode std::sort; rank = position of the group head
//   rounds h = 7, 14, 28, ...: only suffixes in groups larger than one stay; sort them by (rank, rank[i+h]),
//   re-rank, write newly unique suffixes to their final positions.
//
// Output convention of saca() (src/saca.rs:9-15): sa has n+1 entries, sa[0] = n.
#include <parallel/algorithm>
#include <omp.h>
#include <stdint.h>
#include <string.h>

#include <vector>

namespace {
struct Rec {
    uint64_t key;  // round 0: packed symbols; later: (r1 << 32) | r2
    uint32_t idx;
};
inline bool rec_less(const Rec& a, const Rec& b) { return a.key < b.key; }
}  // namespace

extern "C" int oracle_saca_parallel(const uint8_t* s, uint64_t n, uint32_t* sa, int threads) {
    if (threads > 0) omp_set_num_threads(threads);
    sa[0] = (uint32_t)n;
    if (n == 0) return 0;
    const int K = 7;
    std::vector<Rec> rec(n);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        uint64_t key = 0;
        for (int t = 0; t < K; ++t) {
            const uint64_t p = (uint64_t)i + t;
            key = (key << 9) | (p < n ? (uint64_t)s[p] + 1u : 0u);
        }
        rec[i].key = key;
        rec[i].idx = (uint32_t)i;
    }
    __gnu_parallel::sort(rec.begin(), rec.end(), rec_less);
    // rank[i] = SA position (1-based: position 0 is the empty suffix) of the head of i's group; rank[n] = 0
    std::vector<uint32_t> rank(n + 1);
    rank[n] = 0;
    std::vector<uint32_t> head(n);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < (int64_t)n; ++j) head[j] = (j == 0 || rec[j].key != rec[j - 1].key) ? (uint32_t)j : 0u;
    // inclusive prefix maximum (two-pass block scan)
    {
        const int T = omp_get_max_threads();
        std::vector<uint32_t> blockmax(T + 1, 0);
#pragma omp parallel num_threads(T)
        {
            const int t = omp_get_thread_num();
            const uint64_t lo = n * t / T, hi = n * (t + 1) / T;
            uint32_t m = 0;
            for (uint64_t j = lo; j < hi; ++j) {
                if (head[j] > m) m = head[j];
                head[j] = m;
            }
            blockmax[t + 1] = m;
#pragma omp barrier
#pragma omp single
            for (int b = 1; b <= T; ++b)
                if (blockmax[b] < blockmax[b - 1]) blockmax[b] = blockmax[b - 1];
            const uint32_t pre = blockmax[t];
            for (uint64_t j = lo; j < hi; ++j)
                if (head[j] < pre) head[j] = pre;
        }
    }
    std::vector<Rec> act;
    {
        // settled suffixes go to sa[], the others to the active list (r1 in the high word)
        std::vector<uint8_t> keep(n);
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < (int64_t)n; ++j) {
            const uint32_t r = head[j] + 1u;
            rank[rec[j].idx] = r;
            const bool single = (j + 1 >= (int64_t)n || head[j + 1] != head[j]) && head[j] == (uint32_t)j;
            sa[j + 1] = rec[j].idx;
            keep[j] = single ? 0 : 1;
        }
        uint64_t m = 0;
        for (uint64_t j = 0; j < n; ++j) m += keep[j];
        act.resize(m);
        uint64_t w = 0;
        for (uint64_t j = 0; j < n; ++j)
            if (keep[j]) {
                act[w].key = (uint64_t)(head[j] + 1u) << 32;
                act[w].idx = rec[j].idx;
                ++w;
            }
    }
    std::vector<Rec>().swap(rec);
    std::vector<uint32_t>().swap(head);
    uint64_t h = K;
    std::vector<uint32_t> nr;
    while (!act.empty()) {
        const int64_t m = (int64_t)act.size();
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < m; ++j) act[j].key = (act[j].key & 0xffffffff00000000ull) | rank[(uint64_t)act[j].idx + h];
        __gnu_parallel::sort(act.begin(), act.end(), rec_less);
        // new rank = r1 + (index of the new group's head - index of the old group's head)
        nr.assign(m, 0);
        {
            uint32_t ogs = 0, nhs = 0;  // sequential scan: the active list is a small part of the text
            for (int64_t j = 0; j < m; ++j) {
                if (j == 0 || (act[j].key >> 32) != (act[j - 1].key >> 32)) ogs = (uint32_t)j;
                if (j == 0 || act[j].key != act[j - 1].key) nhs = (uint32_t)j;
                nr[j] = (uint32_t)(act[j].key >> 32) + (nhs - ogs);
            }
        }
        std::vector<Rec> next;
        next.reserve(m);
        for (int64_t j = 0; j < m; ++j) {
            const bool first = j == 0 || act[j].key != act[j - 1].key;
            const bool last = j + 1 >= m || act[j + 1].key != act[j].key;
            if (first && last) {
                sa[nr[j]] = act[j].idx;
            } else {
                Rec r;
                r.key = (uint64_t)nr[j] << 32;
                r.idx = act[j].idx;
                next.push_back(r);
            }
        }
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < m; ++j) rank[act[j].idx] = nr[j];
        act.swap(next);
        h *= 2;
        if (h > 2 * n + 16) return -1;
    }
    return 0;
}

extern "C" int oracle_parallel_threads(void) { return omp_get_max_threads(); }
