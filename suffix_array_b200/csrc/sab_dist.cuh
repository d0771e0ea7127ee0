// sab_dist.cuh -- multi-GPU suffix-array construction: one rank per GPU, distributed sample sort of the packed
// keys followed by prefix doubling with a fixed number of exchange steps per round (include/sab200.h,
// "multi-GPU").  The reference has nothing distributed; this is the sharded form of the same
// saca() (/root/reference/src/saca.rs:9-15).
//
//   rank g of P owns text positions [g*B, (g+1)*B) and the rank[] entries the RankLayout gives it
//   1  byte histogram                       all_reduce   -> common code table / key shape (sab_plan_alphabet)
//   2  keys of own positions; splitters     all_gather   of 2048 sampled keys per rank
//   3  partition by destination (onesweep pass, digit = #splitters <= key)         all_to_all (key, i)
//   4  local radix sort -> a contiguous slice of the suffix array; equal keys share a destination, so groups
//      never straddle GPUs and all re-ranking is local
//   5  init_ranks: slice of sa[], active list (r1, i), bucket directory over the slice's sorted keys
//   6  inverse suffix array at the owners: LAZY when at most a quarter of the suffixes is active (only their
//      ranks travel; a request that finds EMPTY is a suffix that was unique after step 4: the owner packs its
//      key from its text shard, the key travels to the slice that holds it and its rank comes back), else all
//      ranks travel and rank[] is dealt block-cyclically (any region of the text is spread over all owners)
//   7  rounds h = k, 2k, ...:  requests i+h -> owners, answers back                2 x all_to_all (+2 lazy)
//      the list keeps its order (answers are placed by list position), so the in-group sort of the
//      single-GPU path applies; re-rank; changed ranks -> owners                     all_to_all
//
// Every step is enqueued on the rank's library stream; the host only waits where it needs a count.
#pragma once
#include <unistd.h>

#include <algorithm>
#include <thread>

#include "sab_comm.cuh"
#include "sab_saca.cuh"

// ------------------------------------------------------------------ kernels
// d_n != null: the record count lives on the device (the grid is sized for the bound n)
template <typename KeyT, typename DigitOp>
__global__ void __launch_bounds__(256) digit_count_kernel(const KeyT* __restrict__ keys, u64 n, const u32* __restrict__ d_n, DigitOp dop,
                                                          u64* __restrict__ counts) {
    SAB_SHARED_ARRAY(u32, s_c, 256);
    s_c[threadIdx.x] = 0;
    if (d_n) n = *d_n;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&s_c[dop(keys[i])], 1u);
    __syncthreads();
    const u32 c = s_c[threadIdx.x];
    if (c) atomicAdd((unsigned long long*)&counts[threadIdx.x], (unsigned long long)c);
}

// One block of 256 threads: gbase[] = exclusive prefix of the 256 bin counts (record offsets of the partition
// pass), and the row this rank contributes to the count exchange: [P counts | nd device-side u32 extras].
__global__ void __launch_bounds__(256)
count_row_kernel(const u64* __restrict__ cnt, u64* __restrict__ gbase, int P, const u32* __restrict__ extra32, int nd,
                 u64* __restrict__ row) {
    SAB_SHARED_ARRAY(u64, s_v, 256);
    const u32 t = threadIdx.x;
    const u64 mine = cnt[t];
    s_v[t] = mine;
    __syncthreads();
    for (u32 off = 1; off < 256; off <<= 1) {
        const u64 v = t >= off ? s_v[t - off] : 0ull;
        __syncthreads();
        s_v[t] += v;
        __syncthreads();
    }
    gbase[t] = s_v[t] - mine;
    if ((int)t < P) row[t] = mine;
    else if ((int)t < P + nd) row[t] = extra32[t - P];
}

__global__ void hist_widen_kernel(const u32* __restrict__ h32, u64* __restrict__ h64) { h64[threadIdx.x] = h32[threadIdx.x]; }

// out[j] = keys[j * step] for the first min(S, ceil(count / step)) samples; out[S] = how many
__global__ void sample_keys_kernel(const u64* __restrict__ keys, u64 count, u64 step, u32 S, u64* __restrict__ out) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    const u64 have = count ? (count + step - 1) / step : 0;
    const u32 ns = have < S ? (u32)have : S;
    if (j < ns) out[j] = keys[(u64)j * step];
    if (j == 0) out[S] = ns;
}

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
// The requests left in owner order; the answers come back in the same order and return to the LIST position
// of their record: key64[pos[t]] = (r1[pos[t]] << 32) | r2[t].  The list stays grouped by r1.
__global__ void dist_place_keys_kernel(const u32* __restrict__ r1, const u32* __restrict__ pos, const u32* __restrict__ r2, u64 m,
                                       u64* __restrict__ key64) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) {
        const u32 p = pos[t];
        key64[p] = ((u64)r1[p] << 32) | r2[t];
    }
}

// SA entries that became final on a GPU that does not hold their position (rebalanced lists): sa[pos] = idx
__global__ void dist_store_sa_kernel(const u32* __restrict__ pos, const u32* __restrict__ idx, u64 n, u32 sa_off,
                                     u32* __restrict__ sa_local) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sa_local[pos[i] - sa_off] = idx[i];
}
// For every cut c = cuts[j] inside the rank-sorted list r1[0..m) (j = blockIdx.x / blocks_per_cut): the smallest
// p in [c + woff, c + woff + window) with r1[p] != r1[c-1] (atomicMin into out[j], pre-set to ~0).
__global__ void find_boundary_kernel(const u32* __restrict__ r1, u64 m, const u64* __restrict__ cuts, u32 blocks_per_cut, u64 woff,
                                     u64 window, unsigned long long* __restrict__ out) {
    const u32 j = blockIdx.x / blocks_per_cut;
    const u64 c = cuts[j];
    if (c == 0 || c >= m) return;
    const u64 t = (u64)(blockIdx.x % blocks_per_cut) * blockDim.x + threadIdx.x;
    const u64 p = c + woff + t;
    if (t < window && p < m && r1[p] != r1[c - 1]) atomicMin(&out[j], (unsigned long long)p);
}

// ---- peer-to-peer rounds: the arenas of all GPUs are mapped into every rank (sab_peers_update), so a SMALL round
// needs no exchange step at all: the gather loads rank[i+h] straight from the owner GPU over NVLink (and resolves
// an EMPTY slot through the owner's text, the slice's directory and sorted keys, all remote loads), the changed
// ranks are stored straight into the owner's block.  Two stream-ordered all_reduces per round order the phases
// (every rank has finished reading before anybody writes, and vice versa) and carry the survivor counts.
struct P2PTable {
    u32* rank_local[SAB_MAX_RANKS];
    const u8* text[SAB_MAX_RANKS];
    const u64* sorted[SAB_MAX_RANKS];
    const u32* dir[SAB_MAX_RANKS];  // pre-shifted: indexed by key >> dir_shift
    u64 R[SAB_MAX_RANKS];
    u64 n_rel[SAB_MAX_RANKS];       // readable text bytes of the shard (own positions + halo)
    u32 sa_off[SAB_MAX_RANKS];
};
__global__ void __launch_bounds__(256)
dist_gather_p2p_kernel(const u32* __restrict__ r1, const u32* __restrict__ idx, u64 m, u32 h, RankLayout lay, P2PTable pt, int lazy,
                       const u16* __restrict__ lut, u32 base, int k, int dir_shift, SplitterDigit sd, u64* __restrict__ key64) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const u64 q = (u64)idx[t] + h;
    const u32 o = lay.owner(q);
    const u64 slot = lay.slot(q, o);
    u32 r2 = pt.rank_local[o][slot];
    if (lazy && r2 == SAB_RANK_EMPTY) {  // block layout: the owner of q holds text position q at offset slot
        const u64 key = pack_key_at(pt.text[o], pt.n_rel[o], lut, base, k, slot);
        const u32 d = sd(key);
        r2 = pt.sa_off[d] + (u32)sorted_key_position(pt.sorted[d], pt.R[d], pt.dir[d], dir_shift, key);
        pt.rank_local[o][slot] = r2;  // memoise; racing writers store the same value
    }
    key64[t] = ((u64)r1[t] << 32) | r2;
}
// rank[idx[t]] = val[t] on the owner of idx[t]; idx == 0xFFFFFFFF marks "nothing to write"
__global__ void __launch_bounds__(256)
dist_scatter_p2p_kernel(const u32* __restrict__ idx, const u32* __restrict__ val, u64 cnt, RankLayout lay, P2PTable pt) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    const u32 i = idx[t];
    if (i == 0xffffffffu) return;
    const u32 o = lay.owner(i);
    pt.rank_local[o][lay.slot(i, o)] = val[t];
}
__global__ void put_u64_kernel(u64* dst, const u32* src32) { *dst = *src32; }

// ---- lazy inverse suffix array, distributed (see the file header, step 6)
__global__ void __launch_bounds__(256)
dist_lazy_collect_kernel(const u32* __restrict__ q, const u32* __restrict__ ans, u64 count, u32 h, u64 shard_lo,
                         const u8* __restrict__ text, u64 n_rel, const u16* __restrict__ lut, u32 base, int k,
                         u64* __restrict__ keys_out, u32* __restrict__ slot_out, u32* __restrict__ counter) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count || ans[t] != SAB_RANK_EMPTY) return;
    const u32 p = atomicAdd(counter, 1u);
    keys_out[p] = pack_key_at(text, n_rel, lut, base, k, (u64)q[t] + h - shard_lo);
    slot_out[p] = (u32)t;
}
// rank of a suffix that was unique after the initial sort = SA position of its key in this slice
__global__ void __launch_bounds__(256)
dist_lookup_kernel(const u64* __restrict__ sorted, u64 R, const u32* __restrict__ dir, int dir_shift, const u64* __restrict__ keys,
                   u64 count, u32 sa_off, u32* __restrict__ rank_out) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    rank_out[t] = sa_off + (u32)sorted_key_position(sorted, R, dir, dir_shift, keys[t]);
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

// ------------------------------------------------------------------ host helpers
static inline unsigned sab_grid(SabContext* c, u64 n, u64 per_block, int waves) {
    u64 g = div_up64(n, per_block);
    const u64 gmax = (u64)c->sm_count * (u64)waves;
    if (g > gmax) g = gmax;
    return (unsigned)(g ? g : 1);
}

// Plans one variable all-to-all with a single host synchronisation:
//   per-destination counts of keys[0..count) (count read from *d_count32 when given; `count` is then its bound)
//   -> exclusive bases of all 256 bins in c->d_gbase (the partition pass of the same records follows)
//   -> row [P counts | nd device u32 extras | nh host u64 extras] all-gathered: rows[s*W + i], W = P + nd + nh
//   -> pl: what this rank sends to / receives from everybody.
template <typename KeyT, typename DigitOp>
static int sab_plan_exchange(sab200_comm* cm, SabContext* c, const KeyT* d_keys, u64 count, const u32* d_count32, DigitOp dop,
                             const u32* d_extra32, int nd, const u64* h_extra, int nh, A2APlan* pl, u64* rows) {
    cudaStream_t st = c->stream;
    const int P = cm->P, W = P + nd + nh;
    u64* d_cnt = c->d_ghist;  // 256 x u64 scratch
    SAB_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 256 * sizeof(u64), st));
    if (count) {
        SAB_LAUNCH((digit_count_kernel<KeyT, DigitOp>), sab_grid(c, count, 256 * 16, 8), 256, 0, st, d_keys, count, d_count32, dop, d_cnt);
        SAB_LAUNCH_CHECK();
        c->stats.kernel_launches++;
    }
    SAB_LAUNCH(count_row_kernel, 1, 256, 0, st, (const u64*)d_cnt, c->d_gbase, P, d_extra32, nd, cm->d_small);
    SAB_LAUNCH_CHECK();
    if (nh) {
        u64* h = cm->h_small + 32;  // pinned staging (rows come back at h_small + 64)
        for (int i = 0; i < nh; ++i) h[i] = h_extra[i];
        SAB_CUDA_TRY(cudaMemcpyAsync(cm->d_small + P + nd, h, (size_t)nh * sizeof(u64), cudaMemcpyHostToDevice, st));
    }
    u64 rows_[SAB_MAX_RANKS * 32];
    u64* rr = rows ? rows : rows_;
    SAB_TRY(sab_comm_gather_rows(cm, st, W, rr));
    pl->stotal = pl->rtotal = 0;
    for (int d = 0; d < P; ++d) {
        pl->scount[d] = rr[(size_t)cm->rank * W + d];
        pl->rcount[d] = rr[(size_t)d * W + cm->rank];
        pl->stotal += pl->scount[d];
        pl->rtotal += pl->rcount[d];
    }
    return SAB_OK;
}

// host mirror of RankLayout::owner / slot (sab_radix.cuh)
static inline u32 sab_layout_owner_host(const RankLayout& L, u64 q) {
    if (L.cyc) return (u32)((q >> L.shift) % (L.pmax + 1u));
    const u32 o = (u32)(q / L.B);
    return o < L.pmax ? o : L.pmax;
}
static inline u64 sab_layout_slot_host(const RankLayout& L, u64 q, u32 o) {
    if (L.cyc) return (((q >> L.shift) / (L.pmax + 1u)) << L.shift) | (q & (u64)(L.B - 1u));
    return q - (u64)o * L.B;
}

// Two-ended bump allocator over the context arena: long-lived arrays grow from the bottom, the temporaries of
// a phase / round from the top (released by restoring `hi`).
struct DistArena {
    char* base;
    size_t lo, hi;
    bool ok;
    template <typename T>
    T* bot(size_t count) {
        const size_t o = sab_align_up(lo, 256), e = o + count * sizeof(T);
        if (e > hi) {
            ok = false;
            return nullptr;
        }
        lo = e;
        return (T*)(base + o);
    }
    template <typename T>
    T* top(size_t count) {
        const size_t bytes = sab_align_up(count * sizeof(T), 256);
        if (hi < lo + bytes) {
            ok = false;
            return nullptr;
        }
        hi -= bytes;
        return (T*)(base + hi);
    }
};
#define SAB_ARENA_CHECK(A)                                                                                     \
    do {                                                                                                       \
        if (!(A).ok) {                                                                                         \
            sab_set_error("%s:%d: device arena exhausted (%zu bytes)", __FILE__, __LINE__, c->arena_bytes);    \
            return SAB_ERR_OOM;                                                                                \
        }                                                                                                      \
    } while (0)

// phase timer: CUDA events at the phase boundaries of the rank's stream, read once at the end
struct DistPhases {
    SabContext* c;
    std::vector<std::pair<int, cudaEvent_t> > ev;
    void mark(int phase) {
        cudaEvent_t e = sab_event_get(c);
        cudaEventRecord(e, c->stream);
        ev.push_back(std::make_pair(phase, e));
    }
    void collect(double* phase_ms, double* total_ms) {  // the stream must be idle
        for (size_t i = 0; i + 1 < ev.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i].second, ev[i + 1].second);
            if (ev[i].first >= 0 && ev[i].first < SAB200_PHASES) phase_ms[ev[i].first] += ms;
        }
        if (ev.size() >= 2) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev.front().second, ev.back().second);
            *total_ms = ms;
        }
        for (auto& p : ev) c->event_pool.push_back(p.second);
        ev.clear();
    }
};

struct DistResult {
    u32* d_slice;
    u64 slice_len, sa_off;
};

static inline int sab_env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Arena for a rank that owns B text positions: sized for slices up to ~1.3 B records with all suffixes active,
// capped by the free device memory (a grow-only arena: no cudaMalloc in steady state).
static inline size_t sab_dist_want(u64 B) {
    const size_t recs = (size_t)(B + B / 3) + ((size_t)1 << 20);
    return recs * 72 + ((size_t)256 << 20);
}
static int sab_dist_reserve(SabContext* c, u64 B) {
    size_t want = sab_dist_want(B);
    if (c->arena_bytes >= want || c->arena_want_seen == want) return SAB_OK;  // steady state: no driver call at all
    c->arena_want_seen = want;
    size_t free_b = 0, total_b = 0;
    SAB_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const size_t avail = (size_t)((double)(free_b + c->arena_bytes) * 0.94);
    if (want > avail) want = avail;
    return sab_arena_reserve(c, want);
}

#ifndef SAB_EMU
// Keeps the peers' arenas mapped into this rank.  rows[s*W + P + {0,1,2,3}] = process id, device, arena address and
// size of rank s (all-gathered with the key counts, so every rank sees the same table and takes the same
// decisions).  A changed entry (first call, re-allocated arena) re-opens the mapping: cudaIpcOpenMemHandle for
// another process, the pointer itself + cudaDeviceEnablePeerAccess inside one process.  Collective.
static int sab_peers_update(sab200_comm* cm, SabContext* c, const u64* rows, int W) {
    const int P = cm->P, g = cm->rank;
    cudaStream_t st = c->stream;
    bool changed = false;
    for (int s = 0; s < P; ++s) {
        const u64* r = rows + (size_t)s * W + P;
        const sab200_comm::Peer& p = cm->peers[s];
        if (p.pid != r[0] || p.dev != r[1] || p.ptr != r[2] || p.bytes != r[3]) changed = true;
    }
    if (!changed) return SAB_OK;  // peers_ok keeps its value
    // 64-byte IPC handles of every arena
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    u64 fail = 0;
    if (cudaIpcGetMemHandle(&mine, c->arena) != cudaSuccess) {
        cudaGetLastError();
        fail = 1;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(cm->h_small + 32, &mine, 64);
    SAB_CUDA_TRY(cudaMemcpyAsync(cm->d_small, cm->h_small + 32, 64, cudaMemcpyHostToDevice, st));
    SAB_TRY(sab_comm_all_gather(cm, st, cm->d_small, cm->d_small + 64, 64));
    SAB_CUDA_TRY(cudaMemcpyAsync(cm->h_small + 64, cm->d_small + 64, (size_t)P * 64, cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    const u64 mypid = rows[(size_t)g * W + P];
    for (int s = 0; s < P; ++s) {
        const u64* r = rows + (size_t)s * W + P;
        sab200_comm::Peer& p = cm->peers[s];
        if (p.ipc && p.base) cudaIpcCloseMemHandle(p.base);
        p.pid = r[0];
        p.dev = r[1];
        p.ptr = r[2];
        p.bytes = r[3];
        p.base = nullptr;
        p.ipc = false;
        if (s == g) {
            p.base = c->arena;
            continue;
        }
        if (p.pid == mypid) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c->device, (int)p.dev) != cudaSuccess || !can) {
                cudaGetLastError();
                fail = 1;
                continue;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess((int)p.dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) fail = 1;
            cudaGetLastError();
            p.base = (void*)(uintptr_t)p.ptr;
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)(cm->h_small + 64) + (size_t)s * 64, 64);
            void* base = nullptr;
            if (cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                fail = 1;
                continue;
            }
            p.base = base;
            p.ipc = true;
        }
    }
    SAB_TRY(sab_comm_sum_u64(cm, st, &fail, 1));
    cm->peers_ok = fail == 0;
    return SAB_OK;
}
#endif

struct DistRun {
    sab200_comm* cm;
    SabContext* c;
    DistArena A;
    DistPhases ph;
    RankLayout lay;
    u32* rank_local;
    u64 lo;  // first text position of this rank

    // rank[idx[t]] = val[t] on the GPU that owns text position idx[t]; records with idx = 0xFFFFFFFF are dropped.
    // d_extra32 / rows: nd device-side u32 values ride along with the count exchange (rows[s*(P+nd) + P + e]).
    int send_ranks(const u32* idx, const u32* val, u64 cnt, const u32* d_extra32 = nullptr, int nd = 0, u64* rows = nullptr) {
        cudaStream_t st = c->stream;
        const size_t mark = A.hi;
        u32* kp = A.top<u32>(cnt + 8);
        u32* vp = A.top<u32>(cnt + 8);
        SAB_ARENA_CHECK(A);
        OwnerDigit dop;
        dop.add = 0;
        dop.lay = lay;
        A2APlan pl;
        SAB_TRY((sab_plan_exchange<u32, OwnerDigit>(cm, c, idx, cnt, nullptr, dop, d_extra32, nd, nullptr, 0, &pl, rows)));
        if (cnt) {
            SAB_TRY(sab_ensure_lookback(c, (size_t)div_up64(cnt, PassShape<u32>::THREADS * PassShape<u32>::ITEMS)));
            SAB_TRY((sab_launch_pass_op<u32, false, OwnerDigit>(c, idx, kp, val, vp, cnt, dop, c->d_gbase)));
        }
        u32* ri = A.top<u32>(pl.rtotal + 8);
        u32* rr = A.top<u32>(pl.rtotal + 8);
        SAB_ARENA_CHECK(A);
        SAB_TRY(sab_comm_exchange(cm, st, pl, kp, ri, sizeof(u32)));
        SAB_TRY(sab_comm_exchange(cm, st, pl, vp, rr, sizeof(u32)));
        if (pl.rtotal) {
            SAB_LAUNCH(dist_scatter_kernel, (unsigned)div_up64(pl.rtotal, 256), 256, 0, st, (const u32*)ri, (const u32*)rr, pl.rtotal,
                       (u32)lo, lay, rank_local);
            SAB_LAUNCH_CHECK();
            c->stats.kernel_launches++;
        }
        A.hi = mark;
        return SAB_OK;
    }
};

// d_text: this rank's shard + halo on the device (text_len bytes).  The arena `A0` starts behind whatever the
// caller placed at its bottom.  Enqueues on c->stream; the stream is idle when the function returns.
static int sab_dist_saca(sab200_comm* cm, SabContext* c, DistArena A0, const u8* d_text, u64 text_len, u64 n, DistResult* res,
                         sab200_dist_stats* ds) {
    const int P = cm->P, g = cm->rank;
    cudaStream_t st = c->stream;
    SabStats& S = c->stats;
    S.n = n;
    cm->bytes_sent = 0;
    cm->collectives = 0;
    const u64 B = n ? div_up64(n, (u64)P) : 1;
    const u64 lo = (u64)g * B < n ? (u64)g * B : n;
    const u64 hi = lo + B < n ? lo + B : n;
    const u64 count = hi - lo;
    res->d_slice = nullptr;
    res->slice_len = 0;
    res->sa_off = 1;
    ds->nranks = (u32)P;
    ds->rank = (u32)g;
    if (n == 0) return SAB_OK;  // the suffix array is the sentinel alone
    const u64 avail = (n - lo) < count + SAB200_SHARD_HALO ? (n - lo) : count + SAB200_SHARD_HALO;
    if (text_len < avail) {
        sab_set_error("rank %d: shard of %llu bytes, need %llu (own positions + halo)", g, (unsigned long long)text_len,
                      (unsigned long long)avail);
        return SAB_ERR_ARGS;
    }
    DistRun R_;
    R_.cm = cm;
    R_.c = c;
    R_.A = A0;
    R_.ph.c = c;
    R_.lo = lo;
    R_.rank_local = nullptr;
    DistArena& A = R_.A;
    DistPhases& ph = R_.ph;
    SAB_TRY(sab_ensure_scan(c, (size_t)div_up64(count + count / 2 + 4096, SAB_GSORT_TILE)));

    // ---- 1. common alphabet / key shape
    ph.mark(0);
    u32* d_h32 = c->d_counters + 16;
    SAB_CUDA_TRY(cudaMemsetAsync(d_h32, 0, 256 * sizeof(u32), st));
    if (count) {
        SAB_LAUNCH(alphabet_hist_kernel, sab_grid(c, count, 256 * 64, 8), 256, 0, st, d_text, count, d_h32);
        SAB_LAUNCH_CHECK();
        S.kernel_launches++;
    }
    SAB_LAUNCH(hist_widen_kernel, 1, 256, 0, st, (const u32*)d_h32, c->d_ghist);
    SAB_LAUNCH_CHECK();
    SAB_TRY(sab_comm_all_reduce_u64(cm, st, c->d_ghist, 256));
    u64* h64 = (u64*)(c->h_small + 1024);
    SAB_CUDA_TRY(cudaMemcpyAsync(h64, c->d_ghist, 256 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    u16 lut[256];
    u32 sigma = 0, base = 2;
    int k = 1, key_bits = 1;
    {
        u64 hh[256];
        memcpy(hh, h64, sizeof(hh));
        sab_plan_alphabet(hh, n, lut, &sigma, &base, &k, &key_bits);
    }
    S.sigma = sigma;
    S.bits_per_symbol = (u32)sab_ceil_log2_u64(base);
    S.symbols_per_key = (u32)k;
    u16* d_lut = (u16*)(c->d_counters + 16 + 256);
    memcpy(c->h_small + 384, lut, sizeof(lut));
    SAB_CUDA_TRY(cudaMemcpyAsync(d_lut, c->h_small + 384, sizeof(lut), cudaMemcpyHostToDevice, st));

    // ---- 2. keys of own positions, splitters from a sample
    ph.mark(1);
    u64* keysA = A.top<u64>(count + 8);
    SAB_ARENA_CHECK(A);
    if (count) {
        sab_prof_begin(c, 2);
        SAB_TRY(sab_launch_pack(c, d_text, avail, count, (const u16*)d_lut, base, k, keysA));
        sab_prof_end(c);
    }
    u64 splitters[SAB_MAX_RANKS];
    for (int i = 0; i < SAB_MAX_RANKS; ++i) splitters[i] = ~0ull;
    if (P > 1) {
        // ~8 Ki samples over all ranks (slice sizes within a few per cent); the pool is sorted on the host
        const u32 NS = 8192 / (u32)P < 512 ? 512u : (8192 / (u32)P > 2048 ? 2048u : 8192 / (u32)P);
        const u64 step = count / NS ? count / NS : 1;
        u64* d_sample = cm->d_small;
        SAB_LAUNCH(sample_keys_kernel, (NS + 255) / 256, 256, 0, st, (const u64*)keysA, count, step, NS, d_sample);
        SAB_LAUNCH_CHECK();
        SAB_TRY(sab_comm_all_gather(cm, st, d_sample, d_sample + (NS + 1), (NS + 1) * sizeof(u64)));
        SAB_CUDA_TRY(cudaMemcpyAsync(cm->h_small, d_sample + (NS + 1), (size_t)P * (NS + 1) * sizeof(u64), cudaMemcpyDeviceToHost, st));
        SAB_CUDA_TRY(cudaStreamSynchronize(st));
        std::vector<u64> pool;
        pool.reserve((size_t)P * NS);
        for (int r = 0; r < P; ++r) {
            const u64* row = cm->h_small + (size_t)r * (NS + 1);
            const u64 have = row[NS] <= NS ? row[NS] : NS;
            pool.insert(pool.end(), row, row + have);
        }
        std::sort(pool.begin(), pool.end());
        for (int i = 0; i + 1 < P; ++i) {
            size_t at = (size_t)(i + 1) * pool.size() / (size_t)P;
            if (at >= pool.size()) at = pool.size() - 1;
            splitters[i] = pool.empty() ? 0 : pool[at];
        }
    }

    // ---- 3. partition by destination (digit = number of splitters <= key), payload = global position
    ph.mark(2);
    SplitterDigit sdop;
    sdop.np = P - 1;
    for (int i = 0; i < SAB_MAX_RANKS - 1; ++i) sdop.s[i] = i < P - 1 ? splitters[i] : ~0ull;
    A2APlan pk;
    u64 mat[SAB_MAX_RANKS * SAB_MAX_RANKS];
    // the count exchange also carries what the peers need to address this rank's receive buffers directly
    // (fused exchange, below): process, device, arena, offset of the key buffer inside it
    const u64 k0_off = sab_align_up(A.lo, 256);
    u64 extra[5] = {(u64)getpid(), (u64)c->device, (u64)(uintptr_t)c->arena, (u64)c->arena_bytes, k0_off};
    constexpr int NX = 5;
    u64 rows[SAB_MAX_RANKS * 32];
    SAB_TRY((sab_plan_exchange<u64, SplitterDigit>(cm, c, keysA, count, nullptr, sdop, nullptr, 0, extra, NX, &pk, rows)));
    const int W = P + NX;
    for (int s = 0; s < P; ++s)
        for (int d = 0; d < P; ++d) mat[(size_t)s * P + d] = rows[(size_t)s * W + d];
    const u64 R = pk.rtotal;
    u64 sa_off = 1;
    for (int r = 0; r < g; ++r)
        for (int s = 0; s < P; ++s) sa_off += mat[(size_t)s * P + r];
    constexpr int PTILE = PassShape<u64>::THREADS * PassShape<u64>::ITEMS;
    SAB_TRY(sab_ensure_lookback(c, (size_t)div_up64((count > R ? count : R) + 1, PTILE)));
    SAB_TRY(sab_ensure_scan(c, (size_t)div_up64(R + 4096, SAB_GSORT_TILE)));
    u64* K0 = A.bot<u64>(R + 8);
    u32* V0 = A.bot<u32>(R + 8);
    SAB_ARENA_CHECK(A);
    bool fused = false;
#ifndef SAB_EMU
    if (P > 1 && cm->kind == 0 && sab_env_int("SAB_DIST_P2P", 1) != 0) {
        SAB_TRY(sab_peers_update(cm, c, rows, W));
        fused = cm->peers_ok;
    }
#endif
    if (fused) {
        // ---- 3+4 fused: the partition pass stores every record straight into its destination GPU's receive
        // buffers over NVLink (rank s writes behind the records of the ranks before it): no staging copy, no
        // all-to-all; the transfer overlaps the ranking tile by tile.  Every rank entered this call (alphabet
        // all_reduce) before anybody stores, and the all_reduce below orders the stores before the local sort.
        PeerOut po;
        memset(&po, 0, sizeof(po));
        u64* h = (u64*)(c->h_small + 1024);
        for (int i = 0; i < 256; ++i) h[i] = 0;
        for (int d = 0; d < P; ++d) {
            u64 Rd = 0, before = 0;
            for (int s = 0; s < P; ++s) {
                Rd += mat[(size_t)s * P + d];
                if (s < g) before += mat[(size_t)s * P + d];
            }
            const u64 koff = rows[(size_t)d * W + P + 4];
            const u64 voff = sab_align_up(koff + (Rd + 8) * sizeof(u64), 256);
            char* base = d == g ? c->arena : (char*)cm->peers[d].base;
            po.k[d] = (u64)(uintptr_t)(base + koff);
            po.v[d] = (u64)(uintptr_t)(base + voff);
            h[d] = before;
        }
        SAB_CUDA_TRY(cudaMemcpyAsync(c->d_gbase, h, 256 * sizeof(u64), cudaMemcpyHostToDevice, st));
        if (count)
            SAB_TRY((sab_launch_pass_op<u64, true, SplitterDigit, true>(c, keysA, (u64*)nullptr, nullptr, (u32*)nullptr, count, sdop,
                                                                       c->d_gbase, &po, (u32)lo)));
        ph.mark(3);
        SAB_TRY(sab_comm_all_reduce_u64(cm, st, cm->d_small, 1));  // barrier: all peers' stores have landed
        for (int d = 0; d < P; ++d)
            if (d != g) cm->bytes_sent += mat[(size_t)g * P + d] * 12;
    } else {
        u64* partK = A.top<u64>(count + 8);
        u32* partI = A.top<u32>(count + 8);
        SAB_ARENA_CHECK(A);
        if (count)
            SAB_TRY((sab_launch_pass_op<u64, true, SplitterDigit>(c, keysA, partK, nullptr, partI, count, sdop, c->d_gbase, nullptr, (u32)lo)));
        // ---- 4. key exchange
        ph.mark(3);
        SAB_TRY(sab_comm_exchange(cm, st, pk, partK, K0, sizeof(u64)));
        SAB_TRY(sab_comm_exchange(cm, st, pk, partI, V0, sizeof(u32)));
    }
    A.hi = A0.hi;  // keysA, partK, partI are dead once the exchange has run (stream order)

    // ---- 5. local sort: this rank's slice of the suffix array
    ph.mark(4);
    SortBuffers<u64> buf;
    buf.k[0] = K0;
    buf.v[0] = V0;
    buf.k[1] = A.bot<u64>(R + 8);
    buf.v[1] = A.bot<u32>(R + 8);
    buf.cur = 0;
    u32* sa_local = A.bot<u32>(R + 8);
    u32* r1buf = A.bot<u32>(R + 8);
    SAB_ARENA_CHECK(A);
    bool sa_written = false;  // the last pass writes the sorted indices straight into the slice of sa[]
    SAB_TRY(sab_radix_sort<u64>(c, buf, R, 0, key_bits, /*iota=*/false, &S.passes[0], sa_local, &sa_written));

    // ---- 6. ranks, slice of sa[], active list, bucket directory over the slice's keys
    ph.mark(5);
    const u64* sortedK = buf.k[buf.cur];
    if (c->want_bkt) SAB_TRY(sab_fused_buckets(c, d_text, n, sortedK, R, (const u16*)d_lut, sigma, base, k));  // this slice's share
    const u32* sortedI = sa_written ? sa_local : buf.v[buf.cur];
    u64* freeK = buf.k[buf.cur ^ 1];
    u32* act_idx = buf.v[buf.cur ^ 1];
    u32* rank_seq = (u32*)freeK;
    int dir_bits = sab_ceil_log2_u64(n) - 4;
    if (dir_bits > 28) dir_bits = 28;
    if (dir_bits > key_bits) dir_bits = key_bits;
    if (dir_bits < 1) dir_bits = 1;
    const int dir_shift = key_bits - dir_bits;
    const u64 kmax = key_bits >= 64 ? ~0ull : ((1ull << key_bits) - 1ull);
    const u64 klo = g > 0 ? splitters[g - 1] : 0ull;
    u64 khi = g < P - 1 ? splitters[g] : kmax;
    if (khi > kmax) khi = kmax;
    if (khi < klo) khi = klo;
    const u64 dlo = klo >> dir_shift;
    const u64 dir_len = (khi >> dir_shift) - dlo + 2;
    u32* dir = A.bot<u32>(dir_len + 8);
    SAB_ARENA_CHECK(A);
    u32* dir_shifted = dir - dlo;  // indexed by key >> dir_shift
    u32* d_m = c->d_counters;
    u64 m = 0;
    if (R) {
        const u64 tiles = div_up64(R, SAB_SCAN_TILE);
        TileState<RankScan> ts = sab_tile_state<RankScan>(c, tiles);
        sab_prof_begin(c, 3);
        SAB_LAUNCH(init_ranks_kernel, (unsigned)tiles, SAB_SCAN_THREADS, 0, st, sortedK, sortedI, R, (u32)sa_off, (u32*)nullptr, rank_seq,
                   sa_written ? (u32*)nullptr : sa_local, r1buf, act_idx, d_m, dir_shifted, dir_shift, ts);
        sab_prof_end(c);
        SAB_LAUNCH_CHECK();
        S.kernel_launches++;
        SAB_CUDA_TRY(cudaMemcpyAsync(c->h_small, d_m, sizeof(u32), cudaMemcpyDeviceToHost, st));
        SAB_CUDA_TRY(cudaStreamSynchronize(st));
        m = c->h_small[0];
    }
    // In lazy mode the round keys ping-pong inside ONE key buffer (the other keeps the sorted keys): the local
    // list must fit its halves.
    const u64 half = sab_align_up((size_t)(R + 8) / 2, 32);
    u64 sums[2] = {m, (m > (R + 8) - half) ? 1ull : 0ull};
    SAB_TRY(sab_comm_sum_u64(cm, st, sums, 2));
    u64 tot = sums[0];
    S.active[0] = m;
    ds->active[0] = tot;
    u32 round = 0;
    bool lazy = false, rebalanced = false;
    int layout_cyc = 0;
    if (tot > 0) {
        // ---- 7. inverse suffix array at the owners
        ph.mark(6);
        lazy = tot <= n / 4 && sums[1] == 0 && sab_env_int("SAB_DIST_LAZY", 1) != 0;
        const char* lenv = getenv("SAB_RANK_LAYOUT");
        layout_cyc = lazy ? 0 : 1;  // lazy look-ups pack keys from the owner's text shard: block ownership
        if (!lazy && lenv && !strcmp(lenv, "block")) layout_cyc = 0;
        u64 local_len;
        if (layout_cyc) {
            u64 per = n / (8ull * (u64)P);
            int shift = 0;
            while (shift < 20 && (2ull << shift) <= per) ++shift;
            const u64 width = 1ull << shift;
            const u64 blocks = div_up64(n + 1, width);
            local_len = div_up64(blocks, (u64)P) << shift;
            sab_rank_layout((u32)width, P, shift, &R_.lay);
        } else {
            local_len = B + 1;  // the last GPU also owns position n
            sab_rank_layout((u32)B, P, -1, &R_.lay);
        }
        const RankLayout lay = R_.lay;
        u32* rank_local = A.bot<u32>(local_len + 8);
        SAB_ARENA_CHECK(A);
        R_.rank_local = rank_local;
        if (lazy) SAB_CUDA_TRY(cudaMemsetAsync(rank_local, 0xff, local_len * sizeof(u32), st));
        {
            const u32 on = sab_layout_owner_host(lay, n);  // the empty suffix has rank 0
            if ((int)on == g) SAB_CUDA_TRY(cudaMemsetAsync(rank_local + sab_layout_slot_host(lay, n, on), 0, sizeof(u32), st));
        }
        if (lazy) SAB_TRY(R_.send_ranks(act_idx, r1buf, m));
        else SAB_TRY(R_.send_ranks(sortedI, rank_seq, R));

        // ---- 8. doubling rounds
        SortBuffers<u64> rb;
        u32 slice_start[SAB_MAX_RANKS];  // first SA position of every slice
        {
            u64 run = 1;
            for (int r = 0; r < SAB_MAX_RANKS; ++r) {
                slice_start[r] = r < P ? (u32)run : 0xffffffffu;
                if (r < P)
                    for (int s = 0; s < P; ++s) run += mat[(size_t)s * P + r];
            }
        }
        u64 key_cap;
        if (lazy) {
            rb.k[0] = freeK;
            rb.k[1] = freeK + half;
            key_cap = (R + 8) - half;
        } else {
            rb.k[0] = freeK;
            rb.k[1] = buf.k[buf.cur];
            key_cap = R + 8;
        }
        rb.v[0] = act_idx;
        rb.v[1] = buf.v[buf.cur];
        rb.cur = 0;
        const u64 val_cap = R + 8;
        u32* sa_shifted = sa_local - (size_t)sa_off;  // ranks are global SA positions
        // ---- 7b. even out the active lists.  The slices hold equal numbers of SUFFIXES, not of active ones (on the
        // mixed text the English-like key ranges hold nearly all of them) and a round costs what its longest list
        // costs.  The lists are globally sorted by rank and groups never straddle GPUs: cut points move to group
        // boundaries and contiguous chunks are shipped, which keeps both properties.  A suffix that becomes unique
        // on a GPU that does not hold its SA position is then routed to the slice owner (step f of a round).
        ph.mark(7);
        if (P > 1) {
            u64 mine[SAB_MAX_RANKS], mm[SAB_MAX_RANKS * SAB_MAX_RANKS];
            for (int d = 0; d < P; ++d) mine[d] = m;
            SAB_TRY(sab_comm_count_matrix(cm, st, mine, mm));  // mm[s*P + *] = list length of rank s
            u64 M = 0, mx = 0, off = 0;
            for (int s = 0; s < P; ++s) {
                const u64 v = mm[(size_t)s * P];
                M += v;
                if (v > mx) mx = v;
                if (s < g) off += v;
            }
            const u64 rebal_min = (u64)sab_env_int("SAB_REBALANCE_MIN", 1 << 20);
            if (M >= rebal_min * (u64)P && (double)mx * P > 1.1 * (double)M) {
                const u64 Q = div_up64(M, (u64)P);
                u64* h = (u64*)(c->h_small + 1024);
                u64 cuts[SAB_MAX_RANKS], found[SAB_MAX_RANKS];
                for (int j = 1; j < P; ++j) {
                    const u64 want = (u64)j * Q;
                    cuts[j - 1] = want <= off ? 0 : (want - off >= m ? m : want - off);
                    found[j - 1] = (cuts[j - 1] == 0 || cuts[j - 1] >= m) ? cuts[j - 1] : ~0ull;
                }
                u64* d_cuts = c->d_ghist;             // [P-1] cuts, then [P-1] results
                unsigned long long* d_found = (unsigned long long*)(c->d_ghist + SAB_MAX_RANKS);
                memcpy(h, cuts, sizeof(u64) * (P - 1));
                SAB_CUDA_TRY(cudaMemcpyAsync(d_cuts, h, sizeof(u64) * (P - 1), cudaMemcpyHostToDevice, st));
                u64 woff = 0, window = 1 << 16;
                for (;;) {
                    bool open = false;
                    for (int j = 0; j < P - 1; ++j) open = open || found[j] == ~0ull;
                    if (!open) break;
                    SAB_CUDA_TRY(cudaMemsetAsync(d_found, 0xff, sizeof(u64) * (P - 1), st));
                    const u32 bpc = (u32)div_up64(window, 256);
                    SAB_LAUNCH(find_boundary_kernel, bpc * (unsigned)(P - 1), 256, 0, st, (const u32*)r1buf, m, (const u64*)d_cuts, bpc, woff,
                               window, d_found);
                    SAB_LAUNCH_CHECK();
                    SAB_CUDA_TRY(cudaMemcpyAsync(h + 32, d_found, sizeof(u64) * (P - 1), cudaMemcpyDeviceToHost, st));
                    SAB_CUDA_TRY(cudaStreamSynchronize(st));
                    for (int j = 0; j < P - 1; ++j) {
                        if (found[j] != ~0ull) continue;
                        if (h[32 + j] != ~0ull) found[j] = h[32 + j];
                        else if (cuts[j] + woff + window >= m) found[j] = m;  // the group runs to the end of the list
                    }
                    woff += window;
                    window *= 8;
                }
                u64 bounds[SAB_MAX_RANKS + 1], sendc[SAB_MAX_RANKS];
                bounds[0] = 0;
                for (int j = 1; j < P; ++j) bounds[j] = found[j - 1] > bounds[j - 1] ? found[j - 1] : bounds[j - 1];
                bounds[P] = m;
                for (int j = 0; j < P; ++j) sendc[j] = bounds[j + 1] - bounds[j];
                A2APlan pb;
                SAB_TRY(sab_comm_plan(cm, st, sendc, &pb, nullptr));
                const u64 cap = key_cap < val_cap ? key_cap : val_cap;
                u64 bad = pb.rtotal + 64 > cap ? 1 : 0;
                SAB_TRY(sab_comm_sum_u64(cm, st, &bad, 1));
                if (bad == 0) {
                    const size_t mark = A.hi;
                    u32* t1 = A.top<u32>(pb.rtotal + 8);
                    u32* t2 = A.top<u32>(pb.rtotal + 8);
                    SAB_ARENA_CHECK(A);
                    SAB_TRY(sab_comm_exchange(cm, st, pb, r1buf, t1, sizeof(u32)));
                    SAB_TRY(sab_comm_exchange(cm, st, pb, rb.v[rb.cur], t2, sizeof(u32)));
                    m = pb.rtotal;
                    if (m) {
                        SAB_CUDA_TRY(cudaMemcpyAsync(r1buf, t1, m * sizeof(u32), cudaMemcpyDeviceToDevice, st));
                        SAB_CUDA_TRY(cudaMemcpyAsync(rb.v[rb.cur], t2, m * sizeof(u32), cudaMemcpyDeviceToDevice, st));
                    }
                    A.hi = mark;
                    rebalanced = true;
                }
            }
        }
        // ---- 7c. peer table of the peer-to-peer rounds (same precondition as the fused key exchange)
        bool p2p_ready = false;
        P2PTable pt;
        memset(&pt, 0, sizeof(pt));
        // records per rank; above it the all-to-all form wins (measured on 8 B200, 1 GiB text: every round peer to peer
        // 20.3 ms, limit 2 Mi 18.8 ms; on 2 GPUs limits of 4 Mi .. 16 Mi are within 1 % of each other)
        const u64 p2p_max = (u64)sab_env_int("SAB_P2P_MAX_RECORDS", 2 << 20);
        if (fused && !rebalanced) {
            // row: offsets of rank_local, text, sorted keys, directory inside the arena; directory origin, slice length, offset
            u64* hrow = cm->h_small + 32;
            hrow[0] = (u64)((char*)rank_local - c->arena);
            hrow[1] = (u64)((const char*)d_text - c->arena);
            hrow[2] = (u64)((const char*)sortedK - c->arena);
            hrow[3] = (u64)((char*)dir - c->arena);
            hrow[4] = dlo;
            hrow[5] = R;
            hrow[6] = sa_off;
            hrow[7] = ((const char*)d_text >= c->arena && (const char*)d_text < c->arena + c->arena_bytes) ? 1 : 0;
            SAB_CUDA_TRY(cudaMemcpyAsync(cm->d_small, hrow, 8 * sizeof(u64), cudaMemcpyHostToDevice, st));
            u64 prow[SAB_MAX_RANKS * 8];
            SAB_TRY(sab_comm_gather_rows(cm, st, 8, prow));  // also orders every rank's ISA scatter before the first gather
            bool ok = true;
            for (int d = 0; d < P; ++d) {
                const u64* r = prow + (size_t)d * 8;
                char* pb = d == g ? c->arena : (char*)cm->peers[d].base;
                ok = ok && r[7] == 1 && pb != nullptr;
                pt.rank_local[d] = (u32*)(pb + r[0]);
                pt.text[d] = (const u8*)(pb + r[1]);
                pt.sorted[d] = (const u64*)(pb + r[2]);
                pt.dir[d] = (const u32*)(pb + r[3]) - r[4];
                pt.R[d] = r[5];
                pt.sa_off[d] = (u32)r[6];
                const u64 lo_d = (u64)d * B < n ? (u64)d * B : n;
                const u64 hi_d = lo_d + B < n ? lo_d + B : n;
                pt.n_rel[d] = (n - lo_d) < (hi_d - lo_d) + SAB200_SHARD_HALO ? (n - lo_d) : (hi_d - lo_d) + SAB200_SHARD_HALO;
            }
            p2p_ready = ok;
        }
        bool lead_barrier = false;  // the first gather is ordered by the row exchange above
        const int rank_bits = sab_ceil_log2_u64(n + 2);
        bool group_sort_on = SAB_GROUP_SORT != 0;
        u64 h = (u64)k;
        while (tot > 0) {
            ++round;
            if (round >= SAB_MAX_ROUNDS || h > n) {
                sab_set_error("prefix doubling did not converge (round %u, h=%llu, active=%llu)", round, (unsigned long long)h,
                              (unsigned long long)tot);
                return SAB_ERR_INTERNAL;
            }
            const size_t topmark = A.hi;
            // Small rounds go peer to peer (no exchange step, one host round trip); large ones through partition +
            // all-to-all (bulk NVLink transfers, local random access).  Same decision on every rank.
            const bool p2p_round = p2p_ready && tot <= p2p_max * (u64)P;
            if (p2p_round) {
                ph.mark(8);
                if (lead_barrier) {  // the owners' scatters of an all-to-all round must have finished everywhere
                    SAB_TRY(sab_comm_all_reduce_u64(cm, st, cm->d_small + 8, 1));
                    lead_barrier = false;
                }
                if (m) {
                    sab_prof_begin(c, 4);
                    SAB_LAUNCH(dist_gather_p2p_kernel, (unsigned)div_up64(m, 256), 256, 0, st, (const u32*)r1buf, (const u32*)rb.v[rb.cur], m,
                               (u32)h, lay, pt, lazy ? 1 : 0, (const u16*)d_lut, base, k, dir_shift, sdop, rb.k[rb.cur]);
                    sab_prof_end(c);
                    SAB_LAUNCH_CHECK();
                    S.kernel_launches++;
                }
            } else {
            // a. requests i + h to the owners (payload = list position), answers back in the same order
            ph.mark(8);
            u32* cur_idx = rb.v[rb.cur];
            u32* ipart = A.top<u32>(m + 8);
            u32* ppart = A.top<u32>(m + 8);
            SAB_ARENA_CHECK(A);
            OwnerDigit odop;
            odop.add = (u32)h;
            odop.lay = lay;
            A2APlan pr;
            SAB_TRY((sab_plan_exchange<u32, OwnerDigit>(cm, c, cur_idx, m, nullptr, odop, nullptr, 0, nullptr, 0, &pr, nullptr)));
            if (m) {
                SAB_TRY(sab_ensure_lookback(c, (size_t)div_up64(m, PTILE)));
                SAB_TRY((sab_launch_pass_op<u32, true, OwnerDigit>(c, cur_idx, ipart, nullptr, ppart, m, odop, c->d_gbase)));
            }
            const u64 nq = pr.rtotal;
            u32* q = A.top<u32>(nq + 8);
            u32* ans = A.top<u32>(nq + 8);
            SAB_ARENA_CHECK(A);
            SAB_TRY(sab_comm_exchange(cm, st, pr, ipart, q, sizeof(u32)));
            if (nq) {
                sab_prof_begin(c, 4);
                SAB_LAUNCH(dist_gather_kernel, (unsigned)div_up64(nq, 256), 256, 0, st, (const u32*)q, nq, (u32)h, (u32)lo, lay,
                           (const u32*)rank_local, ans);
                sab_prof_end(c);
                SAB_LAUNCH_CHECK();
                S.kernel_launches++;
            }
            if (lazy) {
                // b. requests that found EMPTY: key at the owner -> slice that holds the key -> rank back.  Their
                // number stays on the device until it comes back with the count exchange (one synchronisation).
                ph.mark(9);
                u64* keys_u = A.top<u64>(nq + 8);
                u32* slot_u = A.top<u32>(nq + 8);
                SAB_ARENA_CHECK(A);
                u32* d_nu = c->d_counters + 12;
                SAB_CUDA_TRY(cudaMemsetAsync(d_nu, 0, sizeof(u32), st));
                if (nq) {
                    SAB_LAUNCH(dist_lazy_collect_kernel, (unsigned)div_up64(nq, 256), 256, 0, st, (const u32*)q, (const u32*)ans, nq, (u32)h,
                               lo, d_text, avail, (const u16*)d_lut, base, k, keys_u, slot_u, d_nu);
                    SAB_LAUNCH_CHECK();
                    S.kernel_launches++;
                }
                A2APlan pu;
                u64 rows_u[SAB_MAX_RANKS * 32];
                SAB_TRY((sab_plan_exchange<u64, SplitterDigit>(cm, c, keys_u, nq, d_nu, sdop, d_nu, 1, nullptr, 0, &pu, rows_u)));
                const u64 nu = rows_u[(size_t)g * (P + 1) + P];
                ds->resolved_empty += nu;
                u64* kp = A.top<u64>(nu + 8);
                u32* sp = A.top<u32>(nu + 8);
                SAB_ARENA_CHECK(A);
                if (nu) {
                    SAB_TRY(sab_ensure_lookback(c, (size_t)div_up64(nu, PTILE)));
                    SAB_TRY((sab_launch_pass_op<u64, false, SplitterDigit>(c, keys_u, kp, slot_u, sp, nu, sdop, c->d_gbase)));
                }
                u64* kq = A.top<u64>(pu.rtotal + 8);
                u32* rk = A.top<u32>(pu.rtotal + 8);
                u32* back = A.top<u32>(nu + 8);
                SAB_ARENA_CHECK(A);
                SAB_TRY(sab_comm_exchange(cm, st, pu, kp, kq, sizeof(u64)));
                if (pu.rtotal) {
                    SAB_LAUNCH(dist_lookup_kernel, (unsigned)div_up64(pu.rtotal, 256), 256, 0, st, sortedK, R, (const u32*)dir_shifted,
                               dir_shift, (const u64*)kq, pu.rtotal, (u32)sa_off, rk);
                    SAB_LAUNCH_CHECK();
                    S.kernel_launches++;
                }
                SAB_TRY(sab_comm_exchange_back(cm, st, pu, rk, back, sizeof(u32)));
                if (nu) {
                    SAB_LAUNCH(dist_lazy_fill_kernel, (unsigned)div_up64(nu, 256), 256, 0, st, (const u32*)sp, (const u32*)back, nu,
                               (const u32*)q, (u32)h, (u32)lo, ans, rank_local);
                    SAB_LAUNCH_CHECK();
                    S.kernel_launches++;
                }
                ph.mark(8);
            }
            u32* r2 = A.top<u32>(m + 8);
            SAB_ARENA_CHECK(A);
            SAB_TRY(sab_comm_exchange_back(cm, st, pr, ans, r2, sizeof(u32)));
            if (m) {
                SAB_LAUNCH(dist_place_keys_kernel, (unsigned)div_up64(m, 256), 256, 0, st, (const u32*)r1buf, (const u32*)ppart,
                           (const u32*)r2, m, rb.k[rb.cur]);
                SAB_LAUNCH_CHECK();
                S.kernel_launches++;
            }
            }  // all-to-all form of steps a, b
            A.hi = topmark;  // requests, answers and look-up buffers are dead (stream order)

            // c. order inside the groups (as on one GPU: the list is still grouped by r1, ascending)
            ph.mark(10);
            u32* upd_idx = A.top<u32>(m + 8);
            u32* upd_r = A.top<u32>(m + 8);
            u32* set_pos = rebalanced ? A.top<u32>(m + 8) : nullptr;
            const u32* sorted_idx = nullptr;
            int sorted_in = 0;  // which of the two index buffers holds the sorted list
            SAB_ARENA_CHECK(A);
            SAB_CUDA_TRY(cudaMemsetAsync(d_m, 0, sizeof(u32), st));  // kept = 0 unless the re-rank says otherwise
            if (m) {
                SortBuffers<u64> sb;
                sb.k[0] = rb.k[rb.cur];
                sb.k[1] = rb.k[rb.cur ^ 1];
                sb.v[0] = rb.v[rb.cur];
                sb.v[1] = rb.v[rb.cur ^ 1];
                sb.cur = 0;
                int sorted = 0;
                if (group_sort_on) {
                    const u64 used = sab_align_up(m, 64);
                    const u64 cap_keys = key_cap > used ? key_cap - used : 0;
                    const u64 cap_vals = val_cap > used ? val_cap - used : 0;
                    GroupSortSpare sp_;
                    sp_.k[0] = sb.k[0] + used;
                    sp_.k[1] = sb.k[1] + used;
                    sp_.v[0] = sb.v[0] + used;
                    sp_.v[1] = sb.v[1] + used;
                    sp_.pos = r1buf;  // the first ranks live in the key words until the re-rank writes them back
                    sp_.cap = cap_keys < cap_vals ? cap_keys : cap_vals;
                    if (sp_.cap > val_cap) sp_.cap = val_cap;
                    if (sp_.cap * 8 >= m) {
                        u64 nbig = 0;
                        sorted = sab_group_sort(c, sb, m, 32 + rank_bits, sp_, &S.passes[round], &nbig);
                        if (sorted < 0) return sorted;
                        if (nbig * 2 > m) group_sort_on = false;
                    }
                }
                if (!sorted) SAB_TRY(sab_radix_sort<u64>(c, sb, m, 0, 32 + rank_bits, /*iota=*/false, &S.passes[round]));
                // d. re-rank; newly unique suffixes go to sa[], changed ranks are collected for their owners; the
                // number of survivors stays on the device and comes back with the count exchange of step e
                ph.mark(11);
                u32* out_idx = (sb.cur == 0) ? rb.v[rb.cur ^ 1] : rb.v[rb.cur];
                const u64 tiles = div_up64(m, SAB_SCAN_TILE);
                TileState<RerankScan> ts = sab_tile_state<RerankScan>(c, tiles);
                sab_prof_begin(c, 3);
                SAB_LAUNCH(rerank_kernel, (unsigned)tiles, SAB_SCAN_THREADS, 0, st, (const u64*)sb.k[sb.cur], (const u32*)sb.v[sb.cur], m,
                           (u32*)nullptr, sa_shifted, r1buf, out_idx, upd_idx, upd_r, set_pos, d_m, ts);
                sab_prof_end(c);
                SAB_LAUNCH_CHECK();
                S.kernel_launches++;
                sorted_idx = sb.v[sb.cur];
                sorted_in = sb.cur;
            }
            if (rebalanced) {
                // f. new SA entries to the GPUs that hold their positions (before the buffer of the sorted indices
                // is reused for the next list); a rank with an empty list still takes part in the exchange
                ph.mark(13);
                const size_t mark2 = A.hi;
                u32* kp = A.top<u32>(m + 8);
                u32* vp = A.top<u32>(m + 8);
                SAB_ARENA_CHECK(A);
                SliceDigit sld;
                for (int i = 0; i < SAB_MAX_RANKS; ++i) sld.start[i] = slice_start[i];
                sld.pmax = (u32)P - 1;
                A2APlan ps;
                SAB_TRY((sab_plan_exchange<u32, SliceDigit>(cm, c, set_pos, m, nullptr, sld, nullptr, 0, nullptr, 0, &ps, nullptr)));
                if (m) {
                    SAB_TRY(sab_ensure_lookback(c, (size_t)div_up64(m, PTILE)));
                    SAB_TRY((sab_launch_pass_op<u32, false, SliceDigit>(c, set_pos, kp, sorted_idx, vp, m, sld, c->d_gbase)));
                }
                u32* rp = A.top<u32>(ps.rtotal + 8);
                u32* ri = A.top<u32>(ps.rtotal + 8);
                SAB_ARENA_CHECK(A);
                SAB_TRY(sab_comm_exchange(cm, st, ps, kp, rp, sizeof(u32)));
                SAB_TRY(sab_comm_exchange(cm, st, ps, vp, ri, sizeof(u32)));
                if (ps.rtotal) {
                    SAB_LAUNCH(dist_store_sa_kernel, (unsigned)div_up64(ps.rtotal, 256), 256, 0, st, (const u32*)rp, (const u32*)ri,
                               ps.rtotal, (u32)sa_off, sa_local);
                    SAB_LAUNCH_CHECK();
                    S.kernel_launches++;
                }
                A.hi = mark2;
            }
            // e. changed ranks to their owners; the same exchange returns every rank's number of survivors
            ph.mark(12);
            u64 kept = 0;
            if (p2p_round) {
                // all_reduce #1: the survivors add up (termination) and every rank has finished LOADING ranks;
                // then the stores go straight into the owners' blocks; all_reduce #2: every store has landed
                SAB_LAUNCH(put_u64_kernel, 1, 1, 0, st, cm->d_small, (const u32*)d_m);
                SAB_LAUNCH_CHECK();
                SAB_TRY(sab_comm_all_reduce_u64(cm, st, cm->d_small, 1));
                if (m) {
                    SAB_LAUNCH(dist_scatter_p2p_kernel, (unsigned)div_up64(m, 256), 256, 0, st, (const u32*)upd_idx, (const u32*)upd_r, m, lay, pt);
                    SAB_LAUNCH_CHECK();
                    S.kernel_launches++;
                }
                SAB_TRY(sab_comm_all_reduce_u64(cm, st, cm->d_small + 8, 1));
                SAB_CUDA_TRY(cudaMemcpyAsync(cm->h_small, cm->d_small, sizeof(u64), cudaMemcpyDeviceToHost, st));
                SAB_CUDA_TRY(cudaMemcpyAsync(c->h_small, d_m, sizeof(u32), cudaMemcpyDeviceToHost, st));
                SAB_CUDA_TRY(cudaStreamSynchronize(st));
                tot = cm->h_small[0];
                kept = c->h_small[0];
                ds->p2p_rounds++;
            } else {
                u64 rows_k[SAB_MAX_RANKS * 32];
                SAB_TRY(R_.send_ranks(upd_idx, upd_r, m, d_m, 1, rows_k));
                kept = rows_k[(size_t)g * (P + 1) + P];
                tot = 0;
                for (int s2 = 0; s2 < P; ++s2) tot += rows_k[(size_t)s2 * (P + 1) + P];
                lead_barrier = true;
            }
            if (m && sorted_in != 0 && kept > 0)
                SAB_CUDA_TRY(cudaMemcpyAsync(rb.v[rb.cur ^ 1], rb.v[rb.cur], kept * sizeof(u32), cudaMemcpyDeviceToDevice, st));
            A.hi = topmark;
            rb.cur ^= 1;
            m = kept;
            S.active[round] = m;
            ds->active[round] = tot;
            h *= 2;
        }
    }
    ph.mark(-1);
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    ph.collect(ds->phase_ms, &ds->total_ms);
    S.rounds = round;
    S.total_ms = ds->total_ms;
    res->d_slice = sa_local;
    res->slice_len = R;
    res->sa_off = sa_off;
    ds->rounds = round;
    ds->lazy_isa = lazy ? 1u : 0u;
    ds->rank_layout = (u32)layout_cyc;
    ds->rebalanced = rebalanced ? 1u : 0u;
    ds->fused_exchange = fused ? 1u : 0u;
    ds->slice_len = R;
    ds->sa_off = sa_off;
    ds->all_to_all_bytes = cm->bytes_sent;
    ds->collectives = cm->collectives;
    return SAB_OK;
}

// ------------------------------------------------------------------ extern "C": communicators
extern "C" int32_t sab200_comm_unique_id(uint8_t id[128]) {
#ifndef SAB_EMU
    if (!id) return SAB_ERR_ARGS;
    SAB_TRY(sab_nccl_load());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    SAB_NCCL_TRY(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, 128);
    return SAB_OK;
#else
    (void)id;
    sab_set_error("the emulator build has no NCCL");
    return SAB_ERR_NCCL;
#endif
}

extern "C" sab200_comm* sab200_comm_create_nccl(const uint8_t id[128], int32_t rank, int32_t nranks, int32_t device) {
#ifndef SAB_EMU
    if (!id || nranks < 1 || nranks > SAB_MAX_RANKS || rank < 0 || rank >= nranks) {
        sab_set_error("sab200_comm_create_nccl: bad arguments (rank %d of %d)", (int)rank, (int)nranks);
        return nullptr;
    }
    if (sab_nccl_load() != SAB_OK) return nullptr;
    if (!sab_get_context(device)) return nullptr;
    sab200_comm* cm = new (std::nothrow) sab200_comm();
    if (!cm) return nullptr;
    cm->rank = rank;
    cm->P = nranks;
    cm->device = device;
    cm->kind = 0;
    memset(&cm->cb, 0, sizeof(cm->cb));
    memset(&cm->last, 0, sizeof(cm->last));
    if (cudaSetDevice(device) != cudaSuccess || sab_comm_alloc_scratch(cm) != SAB_OK) {
        sab_comm_free(cm);
        return nullptr;
    }
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclComm_t nc = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&nc, nranks, u, rank);
    if (r != ncclSuccess) {
        sab_set_error("ncclCommInitRank(rank %d of %d): %s", (int)rank, (int)nranks, g_nccl.GetErrorString(r));
        sab_comm_free(cm);
        return nullptr;
    }
    cm->nccl = nc;
    return cm;
#else
    (void)id; (void)rank; (void)nranks; (void)device;
    sab_set_error("the emulator build has no NCCL");
    return nullptr;
#endif
}

extern "C" sab200_comm* sab200_comm_create_callbacks(const sab200_comm_callbacks* cb, int32_t rank, int32_t nranks, int32_t device) {
    if (!cb || !cb->all_gather || !cb->all_reduce_sum_u64 || !cb->all_to_all_v || nranks < 1 || nranks > SAB_MAX_RANKS || rank < 0 ||
        rank >= nranks) {
        sab_set_error("sab200_comm_create_callbacks: bad arguments (rank %d of %d)", (int)rank, (int)nranks);
        return nullptr;
    }
    if (!sab_get_context(device)) return nullptr;
    sab200_comm* cm = new (std::nothrow) sab200_comm();
    if (!cm) return nullptr;
    cm->rank = rank;
    cm->P = nranks;
    cm->device = device;
    cm->kind = 1;
    cm->cb = *cb;
    memset(&cm->last, 0, sizeof(cm->last));
    if (sab_comm_alloc_scratch(cm) != SAB_OK) {
        sab_comm_free(cm);
        return nullptr;
    }
    return cm;
}

extern "C" void sab200_comm_destroy(sab200_comm* cm) { sab_comm_free(cm); }

extern "C" int32_t sab200_comm_stats(sab200_comm* cm, sab200_dist_stats* out) {
    if (!cm || !out) return SAB_ERR_ARGS;
    *out = cm->last;
    return SAB_OK;
}

// ------------------------------------------------------------------ one rank of a sharded construction
// host_out_base != null: the slice is copied to host_out_base + sa_off (the caller's whole sa[] array)
// h_bkt_part != null: this rank's share of the fused bucket table (SAB200_BKT_LEN counts, host) -- the caller adds the
// shares of all ranks and 1 for the empty suffix
static int sab_sharded_run(sab200_comm* cm, const u8* shard, u64 shard_len, u64 n, int shard_on_device, u32* out, u64 out_cap,
                           int out_on_device, u32* host_out_base, u64* slice_len, u64* sa_off, const u32** d_slice,
                           u32* h_bkt_part = nullptr) {
    if (!cm || (shard_len > 0 && !shard) || n > SAB200_MAX_LENGTH) {
        sab_set_error("sab200_saca_sharded: bad arguments (n=%llu)", (unsigned long long)n);
        return SAB_ERR_ARGS;
    }
    const double w0 = now_ms();
    SabContext* c = sab_get_context(cm->device);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_CUDA_TRY(cudaSetDevice(c->device));
    memset(&c->stats, 0, sizeof(c->stats));
    c->profiling = g_profiling;
    sab200_dist_stats& ds = cm->last;
    memset(&ds, 0, sizeof(ds));
    const u64 B = n ? div_up64(n, (u64)cm->P) : 1;
    if (cm->last_want != sab_dist_want(B)) {
        // another text size than last time (every rank sees the same change): arenas may be re-allocated, so all
        // peer mappings are closed first, on every rank, before anybody frees
        cm->last_want = sab_dist_want(B);
        sab_comm_close_peers(cm);
        u64 z = 0;
        SAB_TRY(sab_comm_sum_u64(cm, c->stream, &z, 1));
    }
    SAB_TRY(sab_dist_reserve(c, B));
    DistArena A;
    A.base = c->arena;
    A.lo = 0;
    A.hi = c->arena_bytes & ~(size_t)255;
    A.ok = true;
    cudaStream_t st = c->stream;
    const u8* d_text = shard;
    cudaEvent_t e0 = sab_event_get(c), e1 = sab_event_get(c), e2 = sab_event_get(c), e3 = sab_event_get(c);
    cudaEventRecord(e0, st);
    if (!shard_on_device || (cm->P > 1 && cm->kind == 0)) {
        // host shard: upload; device shard on a multi-GPU run: a copy inside the arena, which is what the peers have
        // mapped (the lazy look-ups of the peer-to-peer rounds read the owner's text)
        u8* t = A.bot<u8>((size_t)shard_len + 64);
        SAB_ARENA_CHECK(A);
        if (shard_len && shard_on_device) SAB_CUDA_TRY(cudaMemcpyAsync(t, shard, shard_len, cudaMemcpyDeviceToDevice, st));
        else if (shard_len) SAB_TRY(sab_copy_h2d(c, t, shard, shard_len));
        d_text = t;
    }
    cudaEventRecord(e1, st);
    u32* d_bkt = nullptr;
    if (h_bkt_part) {
        d_bkt = A.bot<u32>((size_t)SAB200_BKT_LEN + 8);
        SAB_ARENA_CHECK(A);
        SAB_CUDA_TRY(cudaMemsetAsync(d_bkt, 0, (size_t)SAB200_BKT_LEN * sizeof(u32), st));
    }
    c->want_bkt = d_bkt;
    c->bkt_add_one = 0;
    DistResult res;
    const double w1 = now_ms();
    int rc = sab_dist_saca(cm, c, A, d_text, shard_len, n, &res, &ds);
    const double w2 = now_ms();
    c->want_bkt = nullptr;
    c->bkt_add_one = 1;
    if (rc != SAB_OK) {
        cudaMemsetAsync(c->d_ticket, 0, sizeof(u32) * 4, st);  // a failed launch may have left the ticket out of step
        cudaStreamSynchronize(st);
        c->ticket_host = 0;
        return rc;
    }
    if (slice_len) *slice_len = res.slice_len;
    if (sa_off) *sa_off = res.sa_off;
    if (d_slice) *d_slice = res.d_slice;
    cudaEventRecord(e2, st);
    if (h_bkt_part) SAB_CUDA_TRY(cudaMemcpyAsync(h_bkt_part, d_bkt, (size_t)SAB200_BKT_LEN * sizeof(u32), cudaMemcpyDeviceToHost, st));
    if (host_out_base) {
        if (res.slice_len) SAB_TRY(sab_copy_d2h(c, host_out_base + res.sa_off, res.d_slice, res.slice_len * sizeof(u32)));
        if (cm->rank == 0) {
            c->h_small[32] = (u32)n;
            memcpy(host_out_base, c->h_small + 32, sizeof(u32));  // sa[0] = n (src/saca.rs:13)
        }
    } else if (out) {
        if (out_cap < res.slice_len) {
            sab_set_error("sab200_saca_sharded: the slice has %llu entries, the buffer holds %llu", (unsigned long long)res.slice_len,
                          (unsigned long long)out_cap);
            return SAB_ERR_ARGS;
        }
        if (res.slice_len && out_on_device)
            SAB_CUDA_TRY(cudaMemcpyAsync(out, res.d_slice, res.slice_len * sizeof(u32), cudaMemcpyDeviceToDevice, st));
        else if (res.slice_len)
            SAB_TRY(sab_copy_d2h(c, out, res.d_slice, res.slice_len * sizeof(u32)));
    }
    cudaEventRecord(e3, st);
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ds.phase_ms[14] = ms;
    cudaEventElapsedTime(&ms, e2, e3);
    ds.phase_ms[15] = ms;
    c->stats.h2d_ms = ds.phase_ms[14];
    c->stats.d2h_ms = ds.phase_ms[15];
    c->event_pool.push_back(e0);
    c->event_pool.push_back(e1);
    c->event_pool.push_back(e2);
    c->event_pool.push_back(e3);
    sab_prof_collect(c);
    g_last_stats = c->stats;
    ds.wall_ms = now_ms() - w0;
    ds.host_setup_ms = w1 - w0;
    ds.host_finish_ms = (w2 - w1) - ds.total_ms;  // host time of the construction that its stream did not cover
    return SAB_OK;
}

extern "C" int32_t sab200_sort_pairs_device(uint64_t* d_k0, uint64_t* d_k1, uint32_t* d_v0, uint32_t* d_v1, uint64_t count,
                                            int32_t key_bits, int32_t device) {
    SabContext* c = sab_get_context(device);
    if (!c) return SAB_ERR_CUDA;
    if (key_bits < 0 || key_bits > 64) return SAB_ERR_ARGS;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_CUDA_TRY(cudaSetDevice(c->device));
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

extern "C" int32_t sab200_group_sort_device(uint64_t* d_k0, uint64_t* d_k1, uint32_t* d_v0, uint32_t* d_v1, uint64_t count,
                                            int32_t key_bits, int32_t ascending, int32_t device, uint64_t* nbig_out) {
    SabContext* c = sab_get_context(device);
    if (!c) return SAB_ERR_CUDA;
    if (key_bits < 32 || key_bits > 64 || count > 0xfffffff0ull) return SAB_ERR_ARGS;
    if (nbig_out) *nbig_out = 0;
    if (count == 0) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_CUDA_TRY(cudaSetDevice(c->device));
    const u64 cap = sab_align_up((size_t)count * 2 + 64, 64);  // room for every record, and for the position sort
    SAB_TRY(sab_arena_reserve(c, (size_t)cap * (2 * 8 + 3 * 4) + 8 * 256));
    c->arena_used = 0;
    GroupSortSpare sp;
    sp.k[0] = sab_arena_take<u64>(c, cap);
    sp.k[1] = sab_arena_take<u64>(c, cap);
    sp.v[0] = sab_arena_take<u32>(c, cap);
    sp.v[1] = sab_arena_take<u32>(c, cap);
    sp.pos = sab_arena_take<u32>(c, cap);
    sp.cap = cap;
    SortBuffers<u64> sb;
    sb.k[0] = d_k0;
    sb.k[1] = d_k1;
    sb.v[0] = d_v0;
    sb.v[1] = d_v1;
    sb.cur = 0;
    u32 passes = 0;
    u64 nbig = 0;
    const int rc = sab_group_sort(c, sb, count, key_bits, sp, &passes, &nbig, ascending != 0);
    if (rc < 0) return rc;
    if (rc == 0) {
        sab_set_error("group sort: the records of large groups did not fit the spare buffers");
        return SAB_ERR_INTERNAL;
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (nbig_out) *nbig_out = nbig;
    return sb.cur;
}

extern "C" int32_t sab200_copy_from_device(void* dst, const void* d_src, uint64_t bytes, int32_t device) {
    if (!dst || !d_src) return bytes ? SAB_ERR_ARGS : SAB_OK;
    SabContext* c = sab_get_context(device);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_CUDA_TRY(cudaSetDevice(c->device));
    SAB_CUDA_TRY(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

extern "C" int32_t sab200_saca_sharded(sab200_comm* cm, const uint8_t* shard, uint64_t shard_len, uint64_t n, int32_t shard_on_device,
                                       uint32_t* out, uint64_t out_cap, int32_t out_on_device, uint64_t* slice_len,
                                       uint64_t* sa_off, const uint32_t** d_slice) {
    return sab_sharded_run(cm, shard, shard_len, n, shard_on_device, out, out_cap, out_on_device, nullptr, slice_len, sa_off, d_slice);
}

// ------------------------------------------------------------------ one process, ngpus GPUs: sab200_saca(..., ngpus > 1)
static std::mutex g_multi_mu;
static std::vector<sab200_comm*> g_multi_comms;  // ncclCommInitAll over devices 0..P-1, cached between calls

static int sab_multi_comms(int P) {
#ifndef SAB_EMU
    if ((int)g_multi_comms.size() == P) return SAB_OK;
    for (auto* cm : g_multi_comms) sab_comm_free(cm);
    g_multi_comms.clear();
    SAB_TRY(sab_nccl_load());
    std::vector<int> devs(P);
    for (int i = 0; i < P; ++i) {
        devs[i] = i;
        if (!sab_get_context(i)) return SAB_ERR_CUDA;
    }
    std::vector<ncclComm_t> nc(P);
    SAB_NCCL_TRY(g_nccl.CommInitAll(nc.data(), P, devs.data()));
    for (int i = 0; i < P; ++i) {
        sab200_comm* cm = new (std::nothrow) sab200_comm();
        if (!cm) return SAB_ERR_OOM;
        cm->rank = i;
        cm->P = P;
        cm->device = i;
        cm->kind = 0;
        cm->nccl = nc[i];
        memset(&cm->cb, 0, sizeof(cm->cb));
        memset(&cm->last, 0, sizeof(cm->last));
        g_multi_comms.push_back(cm);
        SAB_TRY(sab_comm_alloc_scratch(cm));
    }
    return SAB_OK;
#else
    (void)P;
    sab_set_error("the emulator build has no NCCL");
    return SAB_ERR_NCCL;
#endif
}

static int sab_saca_multi(const u8* s, u64 n, u32* sa, int P, u32* bkt) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    SAB_TRY(sab_multi_comms(P));
    std::vector<int> rcs(P, SAB_OK);
    std::vector<std::thread> th;
    std::vector<u32> parts;
    if (bkt) parts.assign((size_t)P * SAB200_BKT_LEN, 0u);
    const u64 B = n ? div_up64(n, (u64)P) : 1;
    for (int g = 0; g < P; ++g) {
        th.emplace_back([&, g]() {
            const u64 lo = (u64)g * B < n ? (u64)g * B : n;
            const u64 hi = lo + B < n ? lo + B : n;
            const u64 len = (n - lo) < (hi - lo) + SAB200_SHARD_HALO ? (n - lo) : (hi - lo) + SAB200_SHARD_HALO;
            rcs[g] = sab_sharded_run(g_multi_comms[g], s + lo, len, n, 0, nullptr, 0, 0, sa, nullptr, nullptr, nullptr,
                                     bkt ? parts.data() + (size_t)g * SAB200_BKT_LEN : nullptr);
        });
    }
    for (auto& t : th) t.join();
    if (n == 0) sa[0] = 0;
    if (bkt) {  // inclusive boundaries: the slices' shares add up; + 1 for the empty suffix (src/sa.rs:98)
        for (u32 i = 0; i < SAB200_BKT_LEN; ++i) {
            u32 v = 1;
            for (int g = 0; g < P; ++g) v += parts[(size_t)g * SAB200_BKT_LEN + i];
            bkt[i] = v;
        }
    }
    for (int g = 0; g < P; ++g)
        if (rcs[g] != SAB_OK) return rcs[g];
    return SAB_OK;
}

extern "C" int32_t sab200_multi_stats(int32_t rank, sab200_dist_stats* out) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    if (!out || rank < 0 || rank >= (int)g_multi_comms.size()) return SAB_ERR_ARGS;
    *out = g_multi_comms[rank]->last;
    return SAB_OK;
}

static void sab_multi_shutdown() {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    for (auto* cm : g_multi_comms) sab_comm_free(cm);
    g_multi_comms.clear();
}
