// sab_scan_kernels.cuh -- the three chained-scan kernels of the construction (ranks after the initial
// sort, re-ranking of a doubling round, split filter) in a warp-striped arrangement.
//
// A tile is SAB_SCAN_THREADS x SAB_SCAN_ITEMS consecutive records; warp w owns a contiguous chunk of
// 32 x ITEMS of them and item k of lane l is record k*32 + l of the chunk, so every global access is a
// full-sector coalesced row.  Flags are turned into per-row ballots: "index of the last head at or before
// me" is a count-leading-zeros on the ballot, "how many kept records before me" a popcount -- no
// shared-memory transposes and one block barrier before the look-back.  Neighbouring records come from
// warp shuffles; only the record just before / after a warp's chunk is fetched separately.
#pragma once
#include "sab_context.cuh"
#include "sab_scan.cuh"

#define SAB_SCAN_THREADS 256
#ifndef SAB_SCAN_ITEMS
#define SAB_SCAN_ITEMS 16
#endif
#define SAB_SCAN_WARPS (SAB_SCAN_THREADS / 32)
#define SAB_SCAN_TILE (SAB_SCAN_THREADS * SAB_SCAN_ITEMS)
#define SAB_WCHUNK (32 * SAB_SCAN_ITEMS)
#ifndef SAB_INIT_MIN_BLOCKS
#define SAB_INIT_MIN_BLOCKS 3
#endif

__device__ __forceinline__ u32 lanemask_le() { return lanemask_lt() | (1u << lane_id()); }
__device__ __forceinline__ u32 high_bit(u32 m) { return 31u - (u32)__clz((int)m); }  // m != 0

// keys of one warp chunk + the records adjacent to it -> previous / next key of every item
struct ChunkKeys {
    u64 key[SAB_SCAN_ITEMS];
    u64 before, after;  // record just before / after the chunk (0 when out of range; callers test indices)
    __device__ __forceinline__ void load(const u64* __restrict__ G, u64 wbase, u64 n) {
        const u32 lane = lane_id();
#pragma unroll
        for (int k = 0; k < SAB_SCAN_ITEMS; ++k) {
            const u64 j = wbase + (u64)k * 32 + lane;
            key[k] = j < n ? G[j] : 0ull;
        }
        u64 b = 0, a = 0;
        if (lane == 0 && wbase > 0 && wbase - 1 < n) b = G[wbase - 1];
        if (lane == 31 && wbase + SAB_WCHUNK < n) a = G[wbase + SAB_WCHUNK];
        before = __shfl_sync(SAB_FULL, b, 0);
        after = __shfl_sync(SAB_FULL, a, 31);
    }
    __device__ __forceinline__ u64 prev(int k) const {
        const u64 up = __shfl_up_sync(SAB_FULL, key[k], 1);
        const u64 edge = k > 0 ? __shfl_sync(SAB_FULL, key[k > 0 ? k - 1 : 0], 31) : before;
        return lane_id() == 0 ? edge : up;
    }
    __device__ __forceinline__ u64 next(int k) const {
        const u64 dn = __shfl_down_sync(SAB_FULL, key[k], 1);
        const u64 edge = k < SAB_SCAN_ITEMS - 1 ? __shfl_sync(SAB_FULL, key[k < SAB_SCAN_ITEMS - 1 ? k + 1 : k], 0) : after;
        return lane_id() == 31 ? edge : dn;
    }
};

// warp aggregates -> (exclusive prefix of this warp inside the tile, tile aggregate); one barrier
template <typename T, typename Op>
__device__ __forceinline__ void warp_aggregates(T mine, Op op, T identity, T& warp_prefix, T& tile_total) {
    SAB_SHARED_ARRAY(T, s_wagg, SAB_SCAN_WARPS);
    if (lane_id() == 0) s_wagg[warp_id()] = mine;
    __syncthreads();
    T pre = identity, tot = identity;
#pragma unroll
    for (int i = 0; i < SAB_SCAN_WARPS; ++i) {
        const T a = s_wagg[i];
        if (i < (int)warp_id()) pre = op(pre, a);
        tot = op(tot, a);
    }
    warp_prefix = pre;
    tile_total = tot;
}

// ------------------------------------------------------------------ 4. ranks after the initial sort
struct RankScan {
    u32 head;  // largest index of a group head seen so far (index 0 is always a head)
    u32 cnt;   // number of active (non-singleton) records seen so far
};
struct RankScanOp {
    __device__ __forceinline__ RankScan operator()(const RankScan& a, const RankScan& b) const {
        RankScan r;
        r.head = a.head > b.head ? a.head : b.head;
        r.cnt = a.cnt + b.cnt;
        return r;
    }
};

// K, I: records sorted by key.  The rank of record j is r = rank_base + (index of the head of j's group)
// (rank_base = SA position of record 0: 1 on a single GPU, the slice offset on a multi-GPU rank).
//   sa_out[j] = I[j]                                  (coalesced copy; sa_out == null: I already IS the slice of the
//                                                      suffix array -- the last radix pass wrote it there -- and
//                                                      only the indices of active records are loaded)
//   records of groups larger than one -> (act_r1, act_idx)
//   rank != null:     rank[I[j]] = r for those active records only (lazy ISA)
//   rank_seq != null: rank_seq[j] = r for every record (multi-GPU: ranks travel to the owner of I[j])
//   dir != null:      dir[key >> dir_shift] = j at the first record of every directory bucket
__global__ void __launch_bounds__(SAB_SCAN_THREADS, SAB_INIT_MIN_BLOCKS)
init_ranks_kernel(const u64* __restrict__ K, const u32* __restrict__ I, u64 n, u32 rank_base, u32* __restrict__ rank,
                  u32* __restrict__ rank_seq, u32* __restrict__ sa_out, u32* __restrict__ act_r1,
                  u32* __restrict__ act_idx, u32* __restrict__ d_count, u32* __restrict__ dir, int dir_shift,
                  TileState<RankScan> st) {
    const u32 tile = blockIdx.x, lane = lane_id();  // 1-D grids are dispatched in block order
    const u64 base = (u64)tile * SAB_SCAN_TILE;
    const u64 wbase = base + (u64)warp_id() * SAB_WCHUNK;
    ChunkKeys ck;
    ck.load(K, wbase, n);
    u32 idx[SAB_SCAN_ITEMS];
    if (sa_out) {
#pragma unroll
        for (int k = 0; k < SAB_SCAN_ITEMS; ++k) {
            const u64 j = wbase + (u64)k * 32 + lane;
            idx[k] = 0;
            if (j < n) {
                idx[k] = I[j];
                sa_out[j] = idx[k];
            }
        }
    }
    u32 hb[SAB_SCAN_ITEMS], ab[SAB_SCAN_ITEMS];
    RankScan mine;
    mine.head = 0;
    mine.cnt = 0;
#pragma unroll
    for (int k = 0; k < SAB_SCAN_ITEMS; ++k) {
        const u64 row = wbase + (u64)k * 32;
        const u64 j = row + lane;
        const u64 pk = ck.prev(k), nk = ck.next(k);
        const bool valid = j < n;
        const bool head = valid && (j == 0 || ck.key[k] != pk);
        const bool next_head = (j + 1 >= n) || nk != ck.key[k];
        const bool active = valid && !(head && next_head);
        hb[k] = __ballot_sync(SAB_FULL, head);
        ab[k] = __ballot_sync(SAB_FULL, active);
        if (dir && head && (j == 0 || (ck.key[k] >> dir_shift) != (pk >> dir_shift))) dir[ck.key[k] >> dir_shift] = (u32)j;
        if (hb[k]) mine.head = (u32)row + high_bit(hb[k]);
        mine.cnt += (u32)__popc(ab[k]);
    }
    if (!sa_out) {  // requested before the look-back so that their latency hides behind it
#pragma unroll
        for (int k = 0; k < SAB_SCAN_ITEMS; ++k) idx[k] = ((ab[k] >> lane) & 1u) ? I[wbase + (u64)k * 32 + lane] : 0u;
    }
    RankScan ident;
    ident.head = 0;
    ident.cnt = 0;
    RankScan wpre, total;
    warp_aggregates<RankScan, RankScanOp>(mine, RankScanOp(), ident, wpre, total);
    const RankScan prefix = tile_exclusive_prefix<RankScan, RankScanOp>(st, tile, total, RankScanOp(), ident);
    RankScan run = RankScanOp()(prefix, wpre);
#pragma unroll
    for (int k = 0; k < SAB_SCAN_ITEMS; ++k) {
        const u64 row = wbase + (u64)k * 32;
        const u64 j = row + lane;
        const u32 m = hb[k] & lanemask_le();
        const u32 r = (m ? (u32)row + high_bit(m) : run.head) + rank_base;
        if (j < n && rank_seq) rank_seq[j] = r;
        if ((ab[k] >> lane) & 1u) {
            const u32 pos = run.cnt + (u32)__popc(ab[k] & lanemask_lt());
            act_r1[pos] = r;
            act_idx[pos] = idx[k];
            if (rank) rank[idx[k]] = r;
        }
        if (hb[k]) run.head = (u32)row + high_bit(hb[k]);
        run.cnt += (u32)__popc(ab[k]);
    }
    if (threadIdx.x == 0 && base + SAB_SCAN_TILE >= n) *d_count = prefix.cnt + total.cnt;  // last tile
}

// ------------------------------------------------------------------ 5b. re-rank + compaction
struct RerankScan {
    u32 ogs;  // index (in the active array) of the head of the old group
    u32 nhs;  // index of the head of the new (refined) group
    u32 cnt;  // records kept (still unsettled) so far
};
struct RerankScanOp {
    __device__ __forceinline__ RerankScan operator()(const RerankScan& a, const RerankScan& b) const {
        RerankScan r;
        r.ogs = a.ogs > b.ogs ? a.ogs : b.ogs;
        r.nhs = a.nhs > b.nhs ? a.nhs : b.nhs;
        r.cnt = a.cnt + b.cnt;
        return r;
    }
};

// S, I: active records sorted by (r1, r2) (S = r1<<32 | r2).  For record j:
//   new_r1 = r1 + (head index of its new group - head index of its old group)
//   changed rank  -> rank[I[j]] = new_r1            (rank != null: single GPU)
//                    upd_idx[j] = I[j], upd_r[j] = new_r1, or upd_idx[j] = 0xFFFFFFFF when unchanged
//                    (upd_idx != null: multi-GPU, the owner of rank[I[j]] is another GPU)
//   singleton     -> sa[new_r1] = I[j] (final), dropped; with set_pos != null the store is left to the
//                    caller instead: set_pos[j] = new_r1 for singletons, 0xFFFFFFFF otherwise (multi-GPU
//                    with rebalanced active lists: SA position new_r1 may belong to another GPU)
//   otherwise     -> appended to (out_r1, out_idx)
__global__ void __launch_bounds__(SAB_SCAN_THREADS, 2)
rerank_kernel(const u64* __restrict__ S, const u32* __restrict__ I, u64 m, u32* __restrict__ rank, u32* __restrict__ sa,
              u32* __restrict__ out_r1, u32* __restrict__ out_idx, u32* __restrict__ upd_idx, u32* __restrict__ upd_r,
              u32* __restrict__ set_pos, u32* __restrict__ d_count, TileState<RerankScan> st) {
    const u32 tile = blockIdx.x, lane = lane_id();
    const u64 base = (u64)tile * SAB_SCAN_TILE;
    const u64 wbase = base + (u64)warp_id() * SAB_WCHUNK;
    ChunkKeys ck;
    ck.load(S, wbase, m);
    u32 idx[SAB_SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SAB_SCAN_ITEMS; ++k) {
        const u64 j = wbase + (u64)k * 32 + lane;
        idx[k] = j < m ? I[j] : 0u;
    }
    u32 ob[SAB_SCAN_ITEMS], nb[SAB_SCAN_ITEMS], kb[SAB_SCAN_ITEMS];
    RerankScan mine;
    mine.ogs = 0;
    mine.nhs = 0;
    mine.cnt = 0;
#pragma unroll
    for (int k = 0; k < SAB_SCAN_ITEMS; ++k) {
        const u64 row = wbase + (u64)k * 32;
        const u64 j = row + lane;
        const u64 pk = ck.prev(k), nk = ck.next(k);
        const bool valid = j < m;
        const bool oldhead = valid && (j == 0 || (u32)(ck.key[k] >> 32) != (u32)(pk >> 32));
        const bool newhead = valid && (j == 0 || ck.key[k] != pk);
        const bool next_newhead = (j + 1 >= m) || nk != ck.key[k];
        const bool keep = valid && !(newhead && next_newhead);
        ob[k] = __ballot_sync(SAB_FULL, oldhead);
        nb[k] = __ballot_sync(SAB_FULL, newhead);
        kb[k] = __ballot_sync(SAB_FULL, keep);
        if (ob[k]) mine.ogs = (u32)row + high_bit(ob[k]);
        if (nb[k]) mine.nhs = (u32)row + high_bit(nb[k]);
        mine.cnt += (u32)__popc(kb[k]);
    }
    RerankScan ident;
    ident.ogs = 0;
    ident.nhs = 0;
    ident.cnt = 0;
    RerankScan wpre, total;
    warp_aggregates<RerankScan, RerankScanOp>(mine, RerankScanOp(), ident, wpre, total);
    const RerankScan prefix = tile_exclusive_prefix<RerankScan, RerankScanOp>(st, tile, total, RerankScanOp(), ident);
    RerankScan run = RerankScanOp()(prefix, wpre);
#pragma unroll
    for (int k = 0; k < SAB_SCAN_ITEMS; ++k) {
        const u64 row = wbase + (u64)k * 32;
        const u64 j = row + lane;
        const u32 le = lanemask_le();
        const u32 mo = ob[k] & le, mn = nb[k] & le;
        const u32 ogs = mo ? (u32)row + high_bit(mo) : run.ogs;
        const u32 nhs = mn ? (u32)row + high_bit(mn) : run.nhs;
        const u32 r1 = (u32)(ck.key[k] >> 32);
        const u32 nr = r1 + (nhs - ogs);
        if (j < m) {
            if (rank && nr != r1) rank[idx[k]] = nr;
            if (upd_idx) {
                upd_idx[j] = nr != r1 ? idx[k] : 0xffffffffu;
                upd_r[j] = nr;
            }
            const bool keep = (kb[k] >> lane) & 1u;
            if (keep) {
                const u32 pos = run.cnt + (u32)__popc(kb[k] & lanemask_lt());
                out_r1[pos] = nr;
                out_idx[pos] = idx[k];
            } else if (!set_pos) {
                sa[nr] = idx[k];
            }
            if (set_pos) set_pos[j] = keep ? 0xffffffffu : nr;
        }
        if (ob[k]) run.ogs = (u32)row + high_bit(ob[k]);
        if (nb[k]) run.nhs = (u32)row + high_bit(nb[k]);
        run.cnt += (u32)__popc(kb[k]);
    }
    if (threadIdx.x == 0 && base + SAB_SCAN_TILE >= m) *d_count = prefix.cnt + total.cnt;
}

// ------------------------------------------------------------------ 5c. skip groups that cannot split
// On repetitive texts most groups gain no information in a round: all their members fetch the same
// second rank.  Such a group keeps its order and its rank, so sorting it is wasted traffic.
// mark_split_groups flags (bitmap over SA positions) the groups that hold two different second ranks;
// split_filter compacts the records of flagged groups in place (they go on to the sort) and moves the
// rest, untouched, straight into the next round's active list.
__global__ void __launch_bounds__(256)
mark_split_groups_kernel(const u64* __restrict__ key64, u64 m, u32* __restrict__ bitmap) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool split = false;
    u32 r1 = 0;
    if (j > 0 && j < m) {
        const u64 a = key64[j - 1], b = key64[j];
        r1 = (u32)(b >> 32);
        split = (a >> 32) == (b >> 32) && (u32)a != (u32)b;
    }
    // At most one atomic per run of neighbouring lanes of a group, and none once the bit is set: in a group of a
    // million records nearly every neighbouring pair differs, and a million atomics on one word serialise in the L2
    // (1 GiB mixed text: 86 ms for this kernel, more than the sort of the round).
    const u32 prev_r1 = __shfl_up_sync(SAB_FULL, r1, 1);
    const u32 prev_split = __shfl_up_sync(SAB_FULL, (u32)split, 1);
    if (split && !(lane_id() > 0 && prev_split && prev_r1 == r1)) {
        const u32 bit = 1u << (r1 & 31u);
        if (!(ld_relaxed_u32(&bitmap[r1 >> 5]) & bit)) atomicOr(&bitmap[r1 >> 5], bit);
    }
}

struct FilterScan {
    u32 sort_cnt;
    u32 stay_cnt;
};
struct FilterScanOp {
    __device__ __forceinline__ FilterScan operator()(const FilterScan& a, const FilterScan& b) const {
        FilterScan r;
        r.sort_cnt = a.sort_cnt + b.sort_cnt;
        r.stay_cnt = a.stay_cnt + b.stay_cnt;
        return r;
    }
};

// key64 / idx_io are compacted IN PLACE: a tile's output range ends before its own input, and the range is only
// written once every predecessor tile has published its aggregate.  A predecessor must therefore have READ its
// input by then: the key loads feed the counts it publishes, the index loads feed nothing before the publication,
// so every thread fences after its loads (a load that is merely issued could still be overtaken by the stores of a
// later tile; ADVICE r1, and an intermittent wrong suffix array on a 12 MiB repetitive text in round 2 -- once in
// about ten runs of tests/test_gpu_parity.py::test_filter_and_group_sort_rounds).
__global__ void __launch_bounds__(SAB_SCAN_THREADS)
split_filter_kernel(u64* key64, u32* idx_io, u64 m, const u32* __restrict__ bitmap, u32* __restrict__ stay_r1,
                    u32* __restrict__ stay_idx, u32* __restrict__ d_counts, TileState<FilterScan> st) {
    const u32 tile = blockIdx.x, lane = lane_id();
    const u64 base = (u64)tile * SAB_SCAN_TILE;
    const u64 wbase = base + (u64)warp_id() * SAB_WCHUNK;
    u64 key[SAB_SCAN_ITEMS];
    u32 idx[SAB_SCAN_ITEMS];
    u32 sb[SAB_SCAN_ITEMS], vb[SAB_SCAN_ITEMS];
    FilterScan mine;
    mine.sort_cnt = 0;
    mine.stay_cnt = 0;
#pragma unroll
    for (int k = 0; k < SAB_SCAN_ITEMS; ++k) {
        const u64 j = wbase + (u64)k * 32 + lane;
        key[k] = 0;
        idx[k] = 0;
        bool flagged = false;
        if (j < m) {
            key[k] = key64[j];
            idx[k] = idx_io[j];
            const u32 r1 = (u32)(key[k] >> 32);
            flagged = (bitmap[r1 >> 5] >> (r1 & 31u)) & 1u;
        }
        sb[k] = __ballot_sync(SAB_FULL, j < m && flagged);
        vb[k] = __ballot_sync(SAB_FULL, j < m && !flagged);
        mine.sort_cnt += (u32)__popc(sb[k]);
        mine.stay_cnt += (u32)__popc(vb[k]);
    }
    fence_acq_rel_gpu();  // the loads above are performed before anything this block publishes (see the comment above)
    FilterScan ident;
    ident.sort_cnt = 0;
    ident.stay_cnt = 0;
    FilterScan wpre, total;
    warp_aggregates<FilterScan, FilterScanOp>(mine, FilterScanOp(), ident, wpre, total);
    const FilterScan prefix = tile_exclusive_prefix<FilterScan, FilterScanOp>(st, tile, total, FilterScanOp(), ident);
    FilterScan run = FilterScanOp()(prefix, wpre);
#pragma unroll
    for (int k = 0; k < SAB_SCAN_ITEMS; ++k) {
        const u32 lt = lanemask_lt();
        if ((sb[k] >> lane) & 1u) {
            const u32 pos = run.sort_cnt + (u32)__popc(sb[k] & lt);
            key64[pos] = key[k];
            idx_io[pos] = idx[k];
        } else if ((vb[k] >> lane) & 1u) {
            const u32 pos = run.stay_cnt + (u32)__popc(vb[k] & lt);
            stay_r1[pos] = (u32)(key[k] >> 32);
            stay_idx[pos] = idx[k];
        }
        run.sort_cnt += (u32)__popc(sb[k]);
        run.stay_cnt += (u32)__popc(vb[k]);
    }
    if (threadIdx.x == 0 && base + SAB_SCAN_TILE >= m) {
        d_counts[0] = prefix.sort_cnt + total.sort_cnt;
        d_counts[1] = prefix.stay_cnt + total.stay_cnt;
    }
}
