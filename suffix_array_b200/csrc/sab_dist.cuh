// sab_dist.cuh -- per-rank step functions of the multi-GPU construction (include/sab200_dist.h).
// The kernels are the single-GPU ones (onesweep with a destination-rank digit, init_ranks / rerank with
// their list outputs); the collectives between the steps belong to the host driver.
#pragma once
#include "../../include/sab200_dist.h"
#include "sab_saca.cuh"

template <typename KeyT, typename DigitOp>
__global__ void __launch_bounds__(256) digit_count_kernel(const KeyT* __restrict__ keys, u64 n, DigitOp dop, u64* __restrict__ counts) {
    SAB_SHARED_ARRAY(u32, s_c, 256);
    s_c[threadIdx.x] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&s_c[dop(keys[i])], 1u);
    __syncthreads();
    const u32 c = s_c[threadIdx.x];
    if (c) atomicAdd((unsigned long long*)&counts[threadIdx.x], (unsigned long long)c);
}

__global__ void iota_base_kernel(u32* __restrict__ out, u64 n, u32 base) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = base + (u32)i;
}
__global__ void hist_widen_kernel(const u32* __restrict__ h32, u64* __restrict__ h64) { h64[threadIdx.x] = h32[threadIdx.x]; }
// local slot of position q on its owner: q - lo under the block layout (lay.cyc = 0), lay.slot() otherwise
__global__ void dist_scatter_kernel(const u32* __restrict__ pos, const u32* __restrict__ val, u64 n, u32 lo, RankLayout lay,
                                    u32* __restrict__ rank_local) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rank_local[lay.cyc ? lay.slot(pos[i], 0) : (u64)(pos[i] - lo)] = val[i];
}
__global__ void dist_gather_kernel(const u32* __restrict__ pos, u64 n, u32 add, u32 lo, RankLayout lay,
                                   const u32* __restrict__ rank_local, u32* __restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const u64 q = (u64)pos[i] + add;
        out[i] = rank_local[lay.cyc ? lay.slot(q, 0) : q - lo];
    }
}
__global__ void dist_make_keys_kernel(const u32* __restrict__ r1, const u32* __restrict__ r2, u64 n, u64* __restrict__ key64) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key64[i] = ((u64)r1[i] << 32) | r2[i];
}

static inline unsigned sab_grid(SabContext* c, u64 n, u64 per_block, int waves) {
    u64 g = div_up64(n, per_block);
    const u64 gmax = (u64)c->sm_count * (u64)waves;
    if (g > gmax) g = gmax;
    return (unsigned)(g ? g : 1);
}

// counts of the first `bins` digits -> host; exclusive prefix -> c->d_gbase[0..256)
template <typename KeyT, typename DigitOp>
static int sab_count_and_base(SabContext* c, const KeyT* d_keys, u64 count, DigitOp dop, int bins, u64* counts_host) {
    cudaStream_t st = c->stream;
    u64* d_cnt = c->d_ghist;  // 256 x u64 scratch
    SAB_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 256 * sizeof(u64), st));
    if (count) {
        SAB_LAUNCH((digit_count_kernel<KeyT, DigitOp>), sab_grid(c, count, 256 * 16, 8), 256, 0, st, d_keys, count, dop, d_cnt);
        SAB_LAUNCH_CHECK();
    }
    u64* h = (u64*)(c->h_small + 1024);  // 512 x u64 of the pinned scratch
    SAB_CUDA_TRY(cudaMemcpyAsync(h, d_cnt, 256 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < bins; ++i) counts_host[i] = h[i];
    u64 run = 0;
    for (int i = 0; i < 256; ++i) {
        const u64 v = h[i];
        h[256 + i] = run;
        run += v;
    }
    SAB_CUDA_TRY(cudaMemcpyAsync(c->d_gbase, h + 256, 256 * sizeof(u64), cudaMemcpyHostToDevice, st));
    return SAB_OK;
}

static SabContext* sab_dist_ctx(int device) {
    SabContext* c = sab_get_context(device);
    if (c) cudaSetDevice(c->device);
    return c;
}

extern "C" int32_t sab200_dist_hist(const uint8_t* d_text, uint64_t len, uint64_t* d_hist, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = c->stream;
    u32* d_h32 = c->d_counters + 16;
    SAB_CUDA_TRY(cudaMemsetAsync(d_h32, 0, 256 * sizeof(u32), st));
    if (len) {
        SAB_LAUNCH(alphabet_hist_kernel, sab_grid(c, len, 256 * 64, 8), 256, 0, st, d_text, len, d_h32);
        SAB_LAUNCH_CHECK();
    }
    SAB_LAUNCH(hist_widen_kernel, 1, 256, 0, st, (const u32*)d_h32, d_hist);
    SAB_LAUNCH_CHECK();
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    return SAB_OK;
}

extern "C" int32_t sab200_dist_plan(const uint64_t* hist256, uint64_t n, uint16_t* lut256, int32_t* b, int32_t* k) {
    if (!hist256 || !lut256 || !b || !k) return SAB_ERR_ARGS;
    u32 sigma = 0, base = 2;
    int kk = 1, bits = 1;
    sab_plan_alphabet(hist256, n, lut256, &sigma, &base, &kk, &bits);
    *b = (int32_t)base;
    *k = kk;
    return SAB_OK;
}

extern "C" int32_t sab200_dist_pack(const uint8_t* d_text, uint64_t shard_lo, uint64_t count, uint64_t n,
                                    const uint16_t* lut256, int32_t b, int32_t k, uint64_t* d_keys, uint32_t* d_idx,
                                    int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    if (shard_lo > n || b < 2 || k < 1 || k > 64) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = c->stream;
    if (count == 0) return SAB_OK;
    u16* d_lut = (u16*)(c->d_counters + 16 + 256);
    memcpy(c->h_small + 384, lut256, 256 * sizeof(u16));
    SAB_CUDA_TRY(cudaMemcpyAsync(d_lut, c->h_small + 384, 256 * sizeof(u16), cudaMemcpyHostToDevice, st));
    // the shard buffer holds count + 64 bytes (halo) at most: never read past them (k <= 64 symbols per key)
    const u64 avail = (n - shard_lo) < count + 64 ? (n - shard_lo) : count + 64;
    SAB_LAUNCH(pack_keys_kernel, (unsigned)div_up64(count, SAB_PACK_TILE), SAB_PACK_THREADS, 0, st, d_text, avail, count,
               (const u16*)d_lut, (u32)b, (int)k, sab_pow_u64((u64)b, k - 1), d_keys);
    SAB_LAUNCH_CHECK();
    SAB_LAUNCH(iota_base_kernel, (unsigned)div_up64(count, 256), 256, 0, st, d_idx, count, (u32)shard_lo);
    SAB_LAUNCH_CHECK();
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    return SAB_OK;
}

extern "C" int32_t sab200_dist_partition_keys(const uint64_t* d_keys, const uint32_t* d_idx, uint64_t count,
                                              const uint64_t* splitters, int32_t nsplit, uint64_t* d_keys_out,
                                              uint32_t* d_idx_out, uint64_t* counts, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    if (nsplit < 0 || nsplit > SAB_MAX_RANKS - 1 || !counts) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    SplitterDigit dop;
    dop.np = nsplit;
    for (int i = 0; i < SAB_MAX_RANKS - 1; ++i) dop.s[i] = i < nsplit ? splitters[i] : ~0ull;
    SAB_TRY((sab_count_and_base<u64, SplitterDigit>(c, d_keys, count, dop, nsplit + 1, counts)));
    if (count) {
        constexpr int TILE = PassShape<u64>::THREADS * PassShape<u64>::ITEMS;
        SAB_TRY(sab_ensure_lookback(c, (size_t)div_up64(count, TILE)));
        SAB_TRY((sab_launch_pass_op<u64, false, SplitterDigit>(c, d_keys, d_keys_out, d_idx, d_idx_out, count, dop, c->d_gbase)));
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

extern "C" int32_t sab200_dist_sort_pairs(uint64_t* d_k0, uint64_t* d_k1, uint32_t* d_v0, uint32_t* d_v1, uint64_t count,
                                          int32_t key_bits, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    if (key_bits < 0 || key_bits > 64) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    SortBuffers<u64> buf;
    buf.k[0] = d_k0;
    buf.k[1] = d_k1;
    buf.v[0] = d_v0;
    buf.v[1] = d_v1;
    buf.cur = 0;
    u32 passes = 0;
    const int rc = sab_radix_sort<u64>(c, buf, count, 0, key_bits, false, &passes);
    if (rc != SAB_OK) return rc;
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return buf.cur;
}

extern "C" int32_t sab200_dist_init_ranks(const uint64_t* d_keys, const uint32_t* d_idx, uint64_t count, uint32_t sa_off,
                                          uint32_t* d_sa_local, uint32_t* d_rank_seq, uint32_t* d_act_r1,
                                          uint32_t* d_act_idx, uint64_t* n_active, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    if (!n_active) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    *n_active = 0;
    if (count == 0) return SAB_OK;
    cudaStream_t st = c->stream;
    const u64 tiles = div_up64(count, SAB_SCAN_TILE);
    SAB_TRY(sab_ensure_scan(c, (size_t)tiles));
    TileState<RankScan> ts = sab_tile_state<RankScan>(c, tiles);
    u32* d_m = c->d_counters;
    SAB_LAUNCH(init_ranks_kernel, (unsigned)tiles, SAB_SCAN_THREADS, 0, st, d_keys, d_idx, count, sa_off, (u32*)nullptr,
               d_rank_seq, d_sa_local, d_act_r1, d_act_idx, d_m, (u32*)nullptr, 0, ts);
    SAB_LAUNCH_CHECK();
    SAB_CUDA_TRY(cudaMemcpyAsync(c->h_small, d_m, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    *n_active = c->h_small[0];
    return SAB_OK;
}

extern "C" int32_t sab200_dist_partition_owner(const uint32_t* d_key, const uint32_t* d_val, uint64_t count, uint32_t add,
                                               uint32_t B, int32_t P, int32_t cyc_shift, uint32_t* d_key_out,
                                               uint32_t* d_val_out, uint64_t* counts, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    OwnerDigit dop;
    dop.add = add;
    if (P > SAB_MAX_RANKS || !counts || sab_rank_layout(B, P, cyc_shift, &dop.lay) != 0) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_TRY((sab_count_and_base<u32, OwnerDigit>(c, d_key, count, dop, P, counts)));
    if (count) {
        constexpr int TILE = PassShape<u32>::THREADS * PassShape<u32>::ITEMS;
        SAB_TRY(sab_ensure_lookback(c, (size_t)div_up64(count, TILE)));
        SAB_TRY((sab_launch_pass_op<u32, false, OwnerDigit>(c, d_key, d_key_out, d_val, d_val_out, count, dop, c->d_gbase)));
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

extern "C" int32_t sab200_dist_partition_slices(const uint32_t* d_pos, const uint32_t* d_val, uint64_t count,
                                                const uint32_t* slice_start, int32_t P, uint32_t* d_pos_out,
                                                uint32_t* d_val_out, uint64_t* counts, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    if (P < 1 || P > SAB_MAX_RANKS || !slice_start || !counts) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    SliceDigit dop;
    for (int i = 0; i < SAB_MAX_RANKS; ++i) dop.start[i] = i < P ? slice_start[i] : 0xffffffffu;
    dop.pmax = (u32)P - 1;
    SAB_TRY((sab_count_and_base<u32, SliceDigit>(c, d_pos, count, dop, P, counts)));
    if (count) {
        constexpr int TILE = PassShape<u32>::THREADS * PassShape<u32>::ITEMS;
        SAB_TRY(sab_ensure_lookback(c, (size_t)div_up64(count, TILE)));
        SAB_TRY((sab_launch_pass_op<u32, false, SliceDigit>(c, d_pos, d_pos_out, d_val, d_val_out, count, dop, c->d_gbase)));
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

extern "C" int32_t sab200_dist_scatter(const uint32_t* d_pos, const uint32_t* d_val, uint64_t count, uint32_t lo, uint32_t B,
                                       int32_t P, int32_t cyc_shift, uint32_t* d_rank_local, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    RankLayout lay;
    if (sab_rank_layout(B, P, cyc_shift, &lay) != 0) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    if (count) {
        SAB_LAUNCH(dist_scatter_kernel, (unsigned)div_up64(count, 256), 256, 0, c->stream, d_pos, d_val, count, lo, lay, d_rank_local);
        SAB_LAUNCH_CHECK();
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

extern "C" int32_t sab200_dist_gather(const uint32_t* d_pos, uint64_t count, uint32_t add, uint32_t lo, uint32_t B, int32_t P,
                                      int32_t cyc_shift, const uint32_t* d_rank_local, uint32_t* d_out, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    RankLayout lay;
    if (sab_rank_layout(B, P, cyc_shift, &lay) != 0) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    if (count) {
        SAB_LAUNCH(dist_gather_kernel, (unsigned)div_up64(count, 256), 256, 0, c->stream, d_pos, count, add, lo, lay, d_rank_local,
                   d_out);
        SAB_LAUNCH_CHECK();
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

extern "C" int32_t sab200_dist_make_keys(const uint32_t* d_r1, const uint32_t* d_r2, uint64_t count, uint64_t* d_key64,
                                         int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    if (count) {
        SAB_LAUNCH(dist_make_keys_kernel, (unsigned)div_up64(count, 256), 256, 0, c->stream, d_r1, d_r2, count, d_key64);
        SAB_LAUNCH_CHECK();
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

extern "C" int32_t sab200_dist_rerank(const uint64_t* d_key64, const uint32_t* d_idx, uint64_t m, uint32_t sa_off,
                                      uint32_t* d_sa_local, uint32_t* d_out_r1, uint32_t* d_out_idx, uint32_t* d_upd_idx,
                                      uint32_t* d_upd_r, uint32_t* d_set_pos, uint64_t* n_kept, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    if (!n_kept) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    *n_kept = 0;
    if (m == 0) return SAB_OK;
    cudaStream_t st = c->stream;
    const u64 tiles = div_up64(m, SAB_SCAN_TILE);
    SAB_TRY(sab_ensure_scan(c, (size_t)tiles));
    TileState<RerankScan> ts = sab_tile_state<RerankScan>(c, tiles);
    u32* d_m = c->d_counters;
    // ranks are global SA positions: index the local slice through a pointer shifted by the slice offset
    u32* sa_shifted = d_sa_local - (size_t)sa_off;
    SAB_LAUNCH(rerank_kernel, (unsigned)tiles, SAB_SCAN_THREADS, 0, st, d_key64, d_idx, m, (u32*)nullptr, sa_shifted, d_out_r1,
               d_out_idx, d_upd_idx, d_upd_r, d_set_pos, d_m, ts);
    SAB_LAUNCH_CHECK();
    SAB_CUDA_TRY(cudaMemcpyAsync(c->h_small, d_m, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    *n_kept = c->h_small[0];
    return SAB_OK;
}

// ------------------------------------------------------------------ lazy inverse suffix array, distributed
// As on one GPU (sab_saca.cuh, step 5) only the ranks of ACTIVE suffixes are stored at their owners; rank[]
// starts EMPTY.  A request that finds EMPTY at the owner concerns a suffix that was unique after the initial
// sort: the owner re-packs its key from its text shard (lazy_collect), the key travels to the GPU whose
// slice holds it (same splitters as the key exchange), that GPU finds it in its sorted keys (lower_bound:
// rank = slice offset + index) and the rank travels back to be stored and answered (lazy_fill).
__global__ void __launch_bounds__(256)
dist_lazy_collect_kernel(const u32* __restrict__ q, const u32* __restrict__ ans, u64 count, u32 h, u64 shard_lo,
                         const u8* __restrict__ text, u64 n_rel, const u16* __restrict__ lut, u32 base, int k,
                         u64* __restrict__ keys_out, u32* __restrict__ slot_out, u32* __restrict__ counter) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count || ans[t] != 0xffffffffu) return;
    const u32 p = atomicAdd(counter, 1u);
    keys_out[p] = pack_key_at(text, n_rel, lut, base, k, (u64)q[t] + h - shard_lo);
    slot_out[p] = (u32)t;
}
__global__ void __launch_bounds__(256)
dist_lower_bound_kernel(const u64* __restrict__ sorted, u64 R, const u64* __restrict__ keys, u64 count, u32 sa_off,
                        u32* __restrict__ rank_out) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const u64 key = keys[t];
    u64 lo = 0, hi = R;
    while (lo < hi) {
        const u64 mid = lo + (hi - lo) / 2;
        if (sorted[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    // the key of a suffix that was unique after the initial sort is present exactly once
    rank_out[t] = (lo < R && sorted[lo] == key) ? sa_off + (u32)lo : 0xffffffffu;
}
__global__ void __launch_bounds__(256)
dist_lazy_fill_kernel(const u32* __restrict__ slot, const u32* __restrict__ rank, u64 count, const u32* __restrict__ q, u32 h,
                      u32 lo, u32* __restrict__ ans, u32* __restrict__ rank_local) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const u32 s = slot[t], r = rank[t];
    ans[s] = r;
    rank_local[(u64)q[s] + h - lo] = r;  // memoised: the next request for this position is answered directly
}

extern "C" int32_t sab200_dist_lazy_collect(const uint32_t* d_q, const uint32_t* d_ans, uint64_t count, uint32_t h,
                                            uint64_t shard_lo, const uint8_t* d_text, uint64_t n, const uint16_t* lut256,
                                            int32_t b, int32_t k, uint64_t* d_keys_out, uint32_t* d_slot_out,
                                            uint64_t* n_unresolved, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    if (!n_unresolved || !lut256 || shard_lo > n || b < 2 || k < 1 || k > 64) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    *n_unresolved = 0;
    if (count == 0) return SAB_OK;
    cudaStream_t st = c->stream;
    u16* d_lut = (u16*)(c->d_counters + 16 + 256);
    memcpy(c->h_small + 384, lut256, 256 * sizeof(u16));
    SAB_CUDA_TRY(cudaMemcpyAsync(d_lut, c->h_small + 384, 256 * sizeof(u16), cudaMemcpyHostToDevice, st));
    u32* d_cnt = c->d_counters + 12;
    SAB_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, sizeof(u32), st));
    SAB_LAUNCH(dist_lazy_collect_kernel, (unsigned)div_up64(count, 256), 256, 0, st, d_q, d_ans, count, h, shard_lo, d_text,
               n - shard_lo, (const u16*)d_lut, (u32)b, (int)k, d_keys_out, d_slot_out, d_cnt);
    SAB_LAUNCH_CHECK();
    SAB_CUDA_TRY(cudaMemcpyAsync(c->h_small + 12, d_cnt, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    *n_unresolved = c->h_small[12];
    return SAB_OK;
}

extern "C" int32_t sab200_dist_lower_bound(const uint64_t* d_sorted_keys, uint64_t R, const uint64_t* d_keys, uint64_t count,
                                           uint32_t sa_off, uint32_t* d_rank_out, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    if (count) {
        SAB_LAUNCH(dist_lower_bound_kernel, (unsigned)div_up64(count, 256), 256, 0, c->stream, d_sorted_keys, R, d_keys, count,
                   sa_off, d_rank_out);
        SAB_LAUNCH_CHECK();
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

extern "C" int32_t sab200_dist_lazy_fill(const uint32_t* d_slot, const uint32_t* d_rank, uint64_t count, const uint32_t* d_q,
                                         uint32_t h, uint32_t lo, uint32_t* d_ans, uint32_t* d_rank_local, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    if (count) {
        SAB_LAUNCH(dist_lazy_fill_kernel, (unsigned)div_up64(count, 256), 256, 0, c->stream, d_slot, d_rank, count, d_q, h, lo,
                   d_ans, d_rank_local);
        SAB_LAUNCH_CHECK();
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

// Bracket a multi-GPU construction on this rank: begin() clears the counters (and arms the per-launch
// events when profiling is on); end() collects them into the record sab200_get_stats() returns.
extern "C" int32_t sab200_dist_begin(int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    memset(&c->stats, 0, sizeof(c->stats));
    c->profiling = g_profiling;
    return SAB_OK;
}
extern "C" int32_t sab200_dist_end(int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    sab_prof_collect(c);
    g_last_stats = c->stats;
    return SAB_OK;
}

// ------------------------------------------------------------------ peer-to-peer rank array (NVLink)
// With the per-rank blocks of rank[] mapped into every process (symmetric memory), a round needs no
// exchange step at all: the gather kernel loads rank[i+h] straight from the owner GPU over NVLink and
// the changed ranks are stored straight into the owner's block.  The host only orders the phases
// (every rank has finished reading before anyone writes, and vice versa).
struct PeerTable {
    u32* p[SAB_MAX_RANKS];
};

// key64[t] = (r1[t] << 32) | rank[idx[t] + h], rank[] distributed over the GPUs as `lay` says
__global__ void __launch_bounds__(256)
dist_gather_p2p_kernel(const u32* __restrict__ r1, const u32* __restrict__ idx, u64 m, u32 h, RankLayout lay, PeerTable pt,
                       u64* __restrict__ key64) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const u64 q = (u64)idx[t] + h;
    const u32 o = lay.owner(q);
    const u32 r2 = pt.p[o][lay.slot(q, o)];
    key64[t] = ((u64)r1[t] << 32) | r2;
}

// rank[idx[t]] = val[t] on the owner of idx[t]; idx == 0xFFFFFFFF marks "nothing to write"
__global__ void __launch_bounds__(256)
dist_scatter_p2p_kernel(const u32* __restrict__ idx, const u32* __restrict__ val, u64 count, RankLayout lay, PeerTable pt) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const u32 i = idx[t];
    if (i == 0xffffffffu) return;
    const u32 o = lay.owner(i);
    pt.p[o][lay.slot(i, o)] = val[t];
}

static int sab_peer_table(const uint64_t* peer_ptrs, int32_t P, PeerTable* pt) {
    if (!peer_ptrs || P < 1 || P > SAB_MAX_RANKS) return SAB_ERR_ARGS;
    for (int i = 0; i < SAB_MAX_RANKS; ++i) pt->p[i] = i < P ? (u32*)(uintptr_t)peer_ptrs[i] : nullptr;
    return SAB_OK;
}

extern "C" int32_t sab200_dist_gather_p2p(const uint32_t* d_r1, const uint32_t* d_idx, uint64_t m, uint32_t h, uint32_t B,
                                          int32_t P, int32_t cyc_shift, const uint64_t* peer_rank_ptrs, uint64_t* d_key64,
                                          int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    PeerTable pt;
    SAB_TRY(sab_peer_table(peer_rank_ptrs, P, &pt));
    RankLayout lay;
    if (sab_rank_layout(B, P, cyc_shift, &lay) != 0) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    if (m) {
        sab_prof_begin(c, 4);
        SAB_LAUNCH(dist_gather_p2p_kernel, (unsigned)div_up64(m, 256), 256, 0, c->stream, d_r1, d_idx, m, h, lay, pt, d_key64);
        sab_prof_end(c);
        SAB_LAUNCH_CHECK();
        c->stats.kernel_launches++;
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

extern "C" int32_t sab200_dist_scatter_p2p(const uint32_t* d_idx, const uint32_t* d_val, uint64_t count, uint32_t B, int32_t P,
                                           int32_t cyc_shift, const uint64_t* peer_rank_ptrs, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    PeerTable pt;
    SAB_TRY(sab_peer_table(peer_rank_ptrs, P, &pt));
    RankLayout lay;
    if (sab_rank_layout(B, P, cyc_shift, &lay) != 0) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    if (count) {
        sab_prof_begin(c, 3);
        SAB_LAUNCH(dist_scatter_p2p_kernel, (unsigned)div_up64(count, 256), 256, 0, c->stream, d_idx, d_val, count, lay, pt);
        sab_prof_end(c);
        SAB_LAUNCH_CHECK();
        c->stats.kernel_launches++;
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

// ------------------------------------------------------------------ key exchange fused into the partition
// counts only (host, P x u64): how many of the keys go to each destination
extern "C" int32_t sab200_dist_count_keys(const uint64_t* d_keys, uint64_t count, const uint64_t* splitters, int32_t nsplit,
                                          uint64_t* counts, int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    if (nsplit < 0 || nsplit > SAB_MAX_RANKS - 1 || !counts) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    SplitterDigit dop;
    dop.np = nsplit;
    for (int i = 0; i < SAB_MAX_RANKS - 1; ++i) dop.s[i] = i < nsplit ? splitters[i] : ~0ull;
    SAB_TRY((sab_count_and_base<u64, SplitterDigit>(c, d_keys, count, dop, nsplit + 1, counts)));
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

// The partition pass writes destination d's records straight into GPU d's receive buffers
// (peer_key_ptrs[d], peer_idx_ptrs[d]: device addresses mapped into this process) starting at record
// offsets[d] (= records the lower ranks send to d): partition and all-to-all in one kernel, the
// transfer overlapping the ranking tile by tile.  The caller barriers before reading the buffers.
extern "C" int32_t sab200_dist_partition_keys_p2p(const uint64_t* d_keys, const uint32_t* d_idx, uint64_t count,
                                                  const uint64_t* splitters, int32_t nsplit, const uint64_t* offsets,
                                                  const uint64_t* peer_key_ptrs, const uint64_t* peer_idx_ptrs,
                                                  int32_t device) {
    SabContext* c = sab_dist_ctx(device);
    if (!c) return SAB_ERR_CUDA;
    if (nsplit < 0 || nsplit > SAB_MAX_RANKS - 1 || !offsets || !peer_key_ptrs || !peer_idx_ptrs) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    if (count == 0) return SAB_OK;
    SplitterDigit dop;
    dop.np = nsplit;
    for (int i = 0; i < SAB_MAX_RANKS - 1; ++i) dop.s[i] = i < nsplit ? splitters[i] : ~0ull;
    PeerOut po;
    memset(&po, 0, sizeof(po));
    u64* h = (u64*)(c->h_small + 1024);
    for (int i = 0; i < 256; ++i) h[i] = 0;
    for (int d = 0; d <= nsplit; ++d) {
        po.k[d] = peer_key_ptrs[d];
        po.v[d] = peer_idx_ptrs[d];
        h[d] = offsets[d];
    }
    SAB_CUDA_TRY(cudaMemcpyAsync(c->d_gbase, h, 256 * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    constexpr int TILE = PassShape<u64>::THREADS * PassShape<u64>::ITEMS;
    SAB_TRY(sab_ensure_lookback(c, (size_t)div_up64(count, TILE)));
    SAB_TRY((sab_launch_pass_op<u64, false, SplitterDigit, true>(c, d_keys, (u64*)nullptr, d_idx, (u32*)nullptr, count, dop,
                                                                  c->d_gbase, &po)));
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}
