// sab_group_sort.cuh -- sorting the records of a doubling round without a full radix sort.
//
// In a round the records arrive grouped by their first rank r1 (the high word of the 64-bit key) and only
// have to be ordered by the second rank INSIDE each group.  Most groups are tiny (on the 1 GiB DNA-like
// text the mean is < 3 records), so eight radix passes over every record (192 B of traffic per record)
// are wasted work:
//
//   group_sort_kernel   one sweep (24 B per record).  A tile is staged in shared memory; every record
//                       finds the bounds of its group by scanning at most SAB_GSORT_MAX neighbours and,
//                       if the group is small and lies inside the tile, takes its slot by counting the
//                       group members that precede it in (key, input position) order -- a stable
//                       counting rank, O(s) shared-memory reads per record.  Records of larger groups, or
//                       of groups cut by a tile border, are compacted in input order (chained scan) into
//                       spare buffers together with their positions.
//   scatter_back_kernel after the radix sort of those "big" records: the j-th sorted record returns to the
//                       j-th recorded position (sorting by (r1, r2) keeps every group in its own range).
//
// No reference counterpart: it replaces part of the work of divsufsort's group refinement
// (third-party crate behind /root/reference/src/saca.rs:14).
#pragma once
#include "sab_scan_kernels.cuh"
#include "sab_sort.cuh"

#ifndef SAB_GSORT_MAX
#define SAB_GSORT_MAX 32  // largest group ordered in shared memory
#endif
#define SAB_GSORT_THREADS 256
#define SAB_GSORT_ITEMS 8
#define SAB_GSORT_TILE (SAB_GSORT_THREADS * SAB_GSORT_ITEMS)
#define SAB_GSORT_SMEM ((SAB_GSORT_TILE + 2) * 8 + SAB_GSORT_TILE * 4 + SAB_GSORT_TILE * 8 + SAB_GSORT_TILE * 4)
#define SAB_GSORT_NOGROUP 0xffffffffu  // never a rank (ranks are <= n <= 2^32 - 2)

struct CountOp {
    __device__ __forceinline__ u32 operator()(u32 a, u32 b) const { return a + b; }
};

__global__ void __launch_bounds__(SAB_GSORT_THREADS)
group_sort_kernel(const u64* __restrict__ kin, const u32* __restrict__ vin, u64 m, u64* __restrict__ kout,
                  u32* __restrict__ vout, u64* __restrict__ big_k, u32* __restrict__ big_v, u32* __restrict__ big_pos,
                  u32 big_cap, u32* __restrict__ d_nbig, TileState<u32> st) {
    SAB_DYN_SMEM(smem);
    u64* s_key = (u64*)smem;                          // [0] record before the tile, [1..TILE] the tile, then the record after
    u64* s_okey = s_key + SAB_GSORT_TILE + 2;         // the tile in output order
    u32* s_val = (u32*)(s_okey + SAB_GSORT_TILE);
    u32* s_oval = s_val + SAB_GSORT_TILE;
    const u32 tile = blockIdx.x, tid = threadIdx.x, lane = lane_id(), w = warp_id();
    const u64 base = (u64)tile * SAB_GSORT_TILE;
    const u32 valid = (m - base < (u64)SAB_GSORT_TILE) ? (u32)(m - base) : (u32)SAB_GSORT_TILE;
    const u64 none = (u64)SAB_GSORT_NOGROUP << 32;
#pragma unroll
    for (int k = 0; k < SAB_GSORT_ITEMS; ++k) {
        const u32 p = tid + k * SAB_GSORT_THREADS;
        if (p < valid) {
            s_key[p + 1] = kin[base + p];
            s_val[p] = vin[base + p];
        }
    }
    if (tid == 0) {
        s_key[0] = base > 0 ? kin[base - 1] : none;
        s_key[valid + 1] = base + valid < m ? kin[base + valid] : none;
    }
    __syncthreads();

    // warp w owns records [w*256, (w+1)*256) of the tile; item k of lane l is record w*256 + k*32 + l
    u32 bigb[SAB_GSORT_ITEMS];
    u32 mine = 0;
#pragma unroll
    for (int k = 0; k < SAB_GSORT_ITEMS; ++k) {
        const u32 p = w * (32 * SAB_GSORT_ITEMS) + k * 32 + lane;
        bool big = false;
        if (p < valid) {
            const u64 key = s_key[p + 1];
            const u32 r1 = (u32)(key >> 32);
            // groups are contiguous: an equal r1 at distance SAB_GSORT_MAX settles "large" with one read, so the
            // neighbour scans below only ever walk over small groups
            if (p >= SAB_GSORT_MAX && (u32)(s_key[p + 1 - SAB_GSORT_MAX] >> 32) == r1) big = true;
            if (p + SAB_GSORT_MAX < valid && (u32)(s_key[p + 1 + SAB_GSORT_MAX] >> 32) == r1) big = true;
            // a = first record of the group (or the scan limit), b = one past its last record
            u32 a = p, b = p + 1;
            if (!big) {
                while (a > 0 && p - a < SAB_GSORT_MAX && (u32)(s_key[a] >> 32) == r1) --a;
                if ((u32)(s_key[a] >> 32) == r1) big = true;  // runs into the previous tile, or longer than the limit
                while (b < valid && b - a <= SAB_GSORT_MAX && (u32)(s_key[b + 1] >> 32) == r1) ++b;
                if (b - a > SAB_GSORT_MAX || (b == valid && (u32)(s_key[valid + 1] >> 32) == r1)) big = true;
            }
            u32 slot = p;
            if (!big) {
                u32 before = 0;
                for (u32 q = a; q < b; ++q) {
                    const u64 o = s_key[q + 1];
                    before += (o < key || (o == key && q < p)) ? 1u : 0u;
                }
                slot = a + before;
            }
            s_okey[slot] = key;  // records of big groups keep their place; the scatter-back overwrites it
            s_oval[slot] = s_val[p];
        }
        bigb[k] = __ballot_sync(SAB_FULL, big);
        mine += (u32)__popc(bigb[k]);
    }
    u32 wpre, total;
    warp_aggregates<u32, CountOp>(mine, CountOp(), 0u, wpre, total);
    const u32 prefix = tile_exclusive_prefix<u32, CountOp>(st, tile, total, CountOp(), 0u);
    u32 run = prefix + wpre;
#pragma unroll
    for (int k = 0; k < SAB_GSORT_ITEMS; ++k) {
        const u32 p = w * (32 * SAB_GSORT_ITEMS) + k * 32 + lane;
        if ((bigb[k] >> lane) & 1u) {
            const u32 j = run + (u32)__popc(bigb[k] & lanemask_lt());
            if (j < big_cap) {
                big_k[j] = s_key[p + 1];
                big_v[j] = s_val[p];
                big_pos[j] = (u32)(base + p);
            }
        }
        run += (u32)__popc(bigb[k]);
    }
    if (tid == 0 && base + SAB_GSORT_TILE >= m) *d_nbig = prefix + total;  // last tile
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SAB_GSORT_ITEMS; ++k) {
        const u32 p = tid + k * SAB_GSORT_THREADS;
        if (p < valid) {
            kout[base + p] = s_okey[p];
            vout[base + p] = s_oval[p];
        }
    }
}

__global__ void __launch_bounds__(256)
scatter_back_kernel(const u64* __restrict__ big_k, const u32* __restrict__ big_v, const u32* __restrict__ big_pos, u32 nbig,
                    u64* __restrict__ kout, u32* __restrict__ vout) {
    const u32 j = blockIdx.x * 256u + threadIdx.x;
    if (j < nbig) {
        const u32 pos = big_pos[j];
        kout[pos] = big_k[j];
        vout[pos] = big_v[j];
    }
}

// Spare memory for the records of big groups: two key and two payload buffers (radix ping-pong) and the
// position list, `cap` records each.
struct GroupSortSpare {
    u64* k[2];
    u32* v[2];
    u32* pos;
    u64 cap;
};

// Sorts the cnt records of (sb.k[0], sb.v[0]) -- grouped by the high key word, the groups in ASCENDING order of
// that word (the scatter-back pairs the j-th radix-sorted large-group record with the j-th recorded position) --
// by the full key.
// Returns 1 with the result in (sb.k[1], sb.v[1]) and sb.cur = 1; returns 0 with the input untouched when
// more records than sp.cap belong to big groups (the caller falls back to the radix sort); < 0 on error.
// *nbig_out = number of records that needed the radix sort.
static int sab_group_sort(SabContext* c, SortBuffers<u64>& sb, u64 cnt, int key_bits, const GroupSortSpare& sp,
                          u32* passes_out, u64* nbig_out) {
    const u64 tiles = div_up64(cnt, SAB_GSORT_TILE);
    SAB_TRY(sab_ensure_scan(c, (size_t)tiles));
    TileState<u32> ts = sab_tile_state<u32>(c, tiles);
    u32* d_nbig = c->d_counters + 10;
    const u32 cap = sp.cap > 0xffffffffull ? 0xffffffffu : (u32)sp.cap;
#ifndef SAB_EMU
    SAB_CUDA_TRY(cudaFuncSetAttribute(group_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SAB_GSORT_SMEM));
#endif
    sab_prof_begin(c, 5);
    SAB_LAUNCH(group_sort_kernel, (unsigned)tiles, SAB_GSORT_THREADS, SAB_GSORT_SMEM, c->stream, (const u64*)sb.k[0],
               (const u32*)sb.v[0], cnt, sb.k[1], sb.v[1], sp.k[0], sp.v[0], sp.pos, cap, d_nbig, ts);
    sab_prof_end(c);
    SAB_LAUNCH_CHECK();
    c->stats.kernel_launches++;
    SAB_CUDA_TRY(cudaMemcpyAsync(c->h_small + 10, d_nbig, sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    const u64 nbig = c->h_small[10];
    c->stats.group_sort_records += cnt;
    c->stats.group_big_records += nbig;
    *nbig_out = nbig;
    if (passes_out) *passes_out = 0;
    if (nbig > sp.cap) return 0;
    if (nbig > 0) {
        SortBuffers<u64> bb;
        bb.k[0] = sp.k[0];
        bb.k[1] = sp.k[1];
        bb.v[0] = sp.v[0];
        bb.v[1] = sp.v[1];
        bb.cur = 0;
        SAB_TRY(sab_radix_sort<u64>(c, bb, nbig, 0, key_bits, /*iota=*/false, passes_out));
        sab_prof_begin(c, 5);
        SAB_LAUNCH(scatter_back_kernel, (unsigned)div_up64(nbig, 256), 256, 0, c->stream, (const u64*)bb.k[bb.cur],
                   (const u32*)bb.v[bb.cur], (const u32*)sp.pos, (u32)nbig, sb.k[1], sb.v[1]);
        sab_prof_end(c);
        SAB_LAUNCH_CHECK();
        c->stats.kernel_launches++;
    }
    sb.cur = 1;
    return 1;
}
