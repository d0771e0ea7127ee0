// sab_api.cu -- context management and the extern "C" surface declared in include/sab200.h.
// Single translation unit: the kernels live in the .cuh headers included below.
#include "../../include/sab200.h"
#include "sab_common.cuh"
#include "sab_context.cuh"
#include "sab_saca.cuh"
#include "sab_search.cuh"
#include "sab_pack.cuh"

#include <chrono>
#include <cstdlib>
#include <new>
#include <thread>

static_assert(sizeof(SabStats) == sizeof(sab200_stats), "SabStats must mirror sab200_stats");
static_assert(SAB_MAX_ROUNDS == SAB200_MAX_ROUNDS, "round table size");

// ------------------------------------------------------------------ error string
static char g_err[1024] = "";
static std::mutex g_err_mu;
void sab_set_error(const char* fmt, ...) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------ contexts
#define SAB_MAX_DEVICES 16
static SabContext* g_ctx[SAB_MAX_DEVICES];
static std::mutex g_ctx_mu;
static bool g_profiling = false;
static SabStats g_last_stats;

int sab_context_init(SabContext* c) {
    SAB_CUDA_TRY(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    SAB_CUDA_TRY(cudaGetDeviceProperties(&prop, c->device));
    c->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    SAB_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SAB_CUDA_TRY(cudaMalloc(&c->d_ghist, sizeof(u64) * SAB_MAX_PASSES * SAB_RADIX_BINS));
    SAB_CUDA_TRY(cudaMalloc(&c->d_gbase, sizeof(u64) * SAB_MAX_PASSES * SAB_RADIX_BINS));
    SAB_CUDA_TRY(cudaMalloc(&c->d_skip, sizeof(u32) * 16));
    SAB_CUDA_TRY(cudaMallocHost(&c->h_small, sizeof(u32) * 4096));
    SAB_CUDA_TRY(cudaMalloc(&c->d_ticket, sizeof(u32) * 4));
    SAB_CUDA_TRY(cudaMemset(c->d_ticket, 0, sizeof(u32) * 4));
    SAB_CUDA_TRY(cudaMalloc(&c->d_counters, sizeof(u32) * 1024));
    SAB_CUDA_TRY(cudaMemset(c->d_counters, 0, sizeof(u32) * 1024));
    c->ticket_host = 0;
    c->lb_epoch = 0;
    c->scan_epoch = 0;
    c->ready = true;
    return SAB_OK;
}

void sab_context_destroy(SabContext* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& e : c->events) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    cudaFree(c->d_ghist);
    cudaFree(c->d_gbase);
    cudaFree(c->d_skip);
    cudaFreeHost(c->h_small);
    for (int i = 0; i < 2; ++i) {
        if (c->bounce[i]) cudaFreeHost(c->bounce[i]);
        if (c->bounce_ev[i]) cudaEventDestroy(c->bounce_ev[i]);
    }
    cudaFree(c->d_lookback);
    cudaFree(c->d_ticket);
    cudaFree(c->d_scan_slots);
    cudaFree(c->d_counters);
    cudaFree(c->arena);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

SabContext* sab_get_context(int device) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (device < 0 || device >= SAB_MAX_DEVICES) {
        sab_set_error("device %d out of range", device);
        return nullptr;
    }
    if (g_ctx[device]) return g_ctx[device];
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= device) {
        sab_set_error("no CUDA device %d available (%s); libsab200 has no CPU fallback", device,
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count too small");
        return nullptr;
    }
    SabContext* c = new (std::nothrow) SabContext();
    if (!c) return nullptr;
    c->device = device;
    if (sab_context_init(c) != SAB_OK) {
        delete c;
        return nullptr;
    }
    g_ctx[device] = c;
    return c;
}

int sab_arena_reserve(SabContext* c, size_t bytes) {
    if (c->arena_bytes >= bytes) return SAB_OK;
    if (c->arena) {
        SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
        SAB_CUDA_TRY(cudaFree(c->arena));
        c->arena = nullptr;
        c->arena_bytes = 0;
    }
    const size_t want = sab_align_up(bytes, (size_t)1 << 21);
    cudaError_t e = cudaMalloc(&c->arena, want);
    if (e != cudaSuccess) {
        c->arena = nullptr;
        sab_set_error("device arena of %zu bytes: %s", want, cudaGetErrorString(e));
        cudaGetLastError();
        return SAB_ERR_OOM;
    }
    c->arena_bytes = want;
    return SAB_OK;
}

// Status words of one radix pass over `tiles` tiles: one per tile and bin, followed by one per group of
// SAB_LB_GROUP tiles and bin (two-level look-back).  lookback_tiles counts allocated 256-word rows.
int sab_ensure_lookback(SabContext* c, size_t tiles) {
    tiles += tiles / (SAB_LB_GROUP > 0 ? SAB_LB_GROUP : 8) + 2;
    if (c->lookback_tiles >= tiles) return SAB_OK;
    if (c->d_lookback) {
        SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
        SAB_CUDA_TRY(cudaFree(c->d_lookback));
        c->d_lookback = nullptr;
        c->lookback_tiles = 0;
    }
    const size_t want = tiles + tiles / 8 + 64;
    SAB_CUDA_TRY(cudaMalloc(&c->d_lookback, want * SAB_RADIX_BINS * sizeof(u64)));
    SAB_CUDA_TRY(cudaMemsetAsync(c->d_lookback, 0, want * SAB_RADIX_BINS * sizeof(u64), c->stream));
    c->lookback_tiles = want;
    c->lb_epoch = 0;
    return SAB_OK;
}

int sab_ensure_scan(SabContext* c, size_t tiles) {
    if (c->scan_tiles >= tiles) return SAB_OK;
    if (c->d_scan_slots) {
        SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_scan_slots);
        c->d_scan_slots = nullptr;
        c->scan_tiles = 0;
    }
    const size_t want = tiles + tiles / 8 + 64;
    SAB_CUDA_TRY(cudaMalloc(&c->d_scan_slots, want * sizeof(ScanSlot)));
    SAB_CUDA_TRY(cudaMemsetAsync(c->d_scan_slots, 0, want * sizeof(ScanSlot), c->stream));
    c->scan_tiles = want;
    c->scan_epoch = 0;
    return SAB_OK;
}

// ------------------------------------------------------------------ per-launch event timing
cudaEvent_t sab_event_get(SabContext* c) {
    if (!c->event_pool.empty()) {
        cudaEvent_t e = c->event_pool.back();
        c->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void sab_prof_begin(SabContext* c, int kind) {
    if (!c->profiling) return;
    SabEventPair p;
    p.a = sab_event_get(c);
    p.b = sab_event_get(c);
    p.kind = kind;
    cudaEventRecord(p.a, c->stream);
    c->events.push_back(p);
}
void sab_prof_end(SabContext* c) {
    if (!c->profiling || c->events.empty()) return;
    cudaEventRecord(c->events.back().b, c->stream);
}
void sab_prof_collect(SabContext* c) {
    if (!c->profiling) return;
    cudaStreamSynchronize(c->stream);
    for (auto& p : c->events) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, p.a, p.b);
        switch (p.kind) {
            case 0: c->stats.radix_pass_ms += ms; break;
            case 1: c->stats.hist_ms += ms; break;
            case 2: c->stats.pack_ms += ms; break;
            case 3: c->stats.rank_ms += ms; break;
            case 4: c->stats.gather_ms += ms; break;
            case 5: c->stats.group_sort_ms += ms; break;
        }
        c->event_pool.push_back(p.a);
        c->event_pool.push_back(p.b);
    }
    c->events.clear();
}

// ------------------------------------------------------------------ host <-> device copies of caller buffers
// The seam hands over whatever the caller allocated -- for the Rust crate a plain Vec (pageable memory).  CUDA
// stages pageable copies through a small internal buffer at 12-15 GB/s; here they go through two pinned bounce
// buffers of the context instead: the DMA of chunk i+1 (PCIe rate) overlaps the host memcpy of chunk i, which a
// few threads share.  Pinned / registered caller buffers take the direct path.
#define SAB_BOUNCE_BYTES ((size_t)32 << 20)
static int sab_copy_threads() {
    static int t = 0;
    if (!t) {
        int hw = (int)std::thread::hardware_concurrency();
        t = hw >= 16 ? 6 : (hw >= 8 ? 4 : (hw >= 4 ? 2 : 1));
        const char* e = getenv("SAB_COPY_THREADS");
        if (e && atoi(e) > 0) t = atoi(e);
    }
    return t;
}
static void sab_parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const int T = sab_copy_threads();
    if (T <= 1 || bytes < ((size_t)4 << 20)) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (((bytes + (size_t)T - 1) / (size_t)T) + 4095) & ~(size_t)4095;  // T * per >= bytes
    for (int i = 1; i < T; ++i) {
        const size_t lo = (size_t)i * per;
        if (lo >= bytes) break;
        const size_t len = lo + per < bytes ? per : bytes - lo;
        th.emplace_back([=]() { memcpy((char*)dst + lo, (const char*)src + lo, len); });
    }
    memcpy(dst, src, per < bytes ? per : bytes);
    for (auto& t : th) t.join();
}
static bool sab_is_pageable(const void* p) {
    const char* f = getenv("SAB_FORCE_STAGED");
    if (f && *f == '1') return true;
#ifdef SAB_EMU
    (void)p;
    return false;
#else
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
#endif
}
static int sab_bounce_init(SabContext* c) {
    if (c->bounce[0]) return SAB_OK;
    for (int i = 0; i < 2; ++i) {
        SAB_CUDA_TRY(cudaMallocHost(&c->bounce[i], SAB_BOUNCE_BYTES));
        SAB_CUDA_TRY(cudaEventCreateWithFlags(&c->bounce_ev[i], cudaEventDisableTiming));
    }
    return SAB_OK;
}
// host -> device on c->stream; returns with the copy ENQUEUED (pinned source) or COMPLETE (pageable source)
int sab_copy_h2d(SabContext* c, void* d_dst, const void* h_src, size_t bytes) {
    if (!bytes) return SAB_OK;
    if (!sab_is_pageable(h_src)) {
        SAB_CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c->stream));
        return SAB_OK;
    }
    SAB_TRY(sab_bounce_init(c));
    size_t off = 0;
    for (int i = 0; off < bytes; ++i) {
        const int b = i & 1;
        const size_t len = bytes - off < SAB_BOUNCE_BYTES ? bytes - off : SAB_BOUNCE_BYTES;
        if (i >= 2) SAB_CUDA_TRY(cudaEventSynchronize(c->bounce_ev[b]));  // the DMA out of this buffer has finished
        sab_parallel_memcpy(c->bounce[b], (const char*)h_src + off, len);
        SAB_CUDA_TRY(cudaMemcpyAsync((char*)d_dst + off, c->bounce[b], len, cudaMemcpyHostToDevice, c->stream));
        SAB_CUDA_TRY(cudaEventRecord(c->bounce_ev[b], c->stream));
        off += len;
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}
// device -> host; everything enqueued on c->stream before the call is waited for; returns with the data in place
int sab_copy_d2h(SabContext* c, void* h_dst, const void* d_src, size_t bytes) {
    if (!bytes) return SAB_OK;
    if (!sab_is_pageable(h_dst)) {
        SAB_CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
        SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
        return SAB_OK;
    }
    SAB_TRY(sab_bounce_init(c));
    const size_t nchunks = (bytes + SAB_BOUNCE_BYTES - 1) / SAB_BOUNCE_BYTES;
    auto issue = [&](size_t i) -> int {
        const size_t off = i * SAB_BOUNCE_BYTES;
        const size_t len = bytes - off < SAB_BOUNCE_BYTES ? bytes - off : SAB_BOUNCE_BYTES;
        SAB_CUDA_TRY(cudaMemcpyAsync(c->bounce[i & 1], (const char*)d_src + off, len, cudaMemcpyDeviceToHost, c->stream));
        SAB_CUDA_TRY(cudaEventRecord(c->bounce_ev[i & 1], c->stream));
        return SAB_OK;
    };
    SAB_TRY(issue(0));
    for (size_t i = 0; i < nchunks; ++i) {
        if (i + 1 < nchunks) SAB_TRY(issue(i + 1));  // its buffer was emptied by the memcpy of chunk i-1
        SAB_CUDA_TRY(cudaEventSynchronize(c->bounce_ev[i & 1]));
        const size_t off = i * SAB_BOUNCE_BYTES;
        const size_t len = bytes - off < SAB_BOUNCE_BYTES ? bytes - off : SAB_BOUNCE_BYTES;
        sab_parallel_memcpy((char*)h_dst + off, c->bounce[i & 1], len);
    }
    return SAB_OK;
}

static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

static int run_device(SabContext* c, const u8* d_s, u64 n, u32* d_sa) {
    c->profiling = g_profiling;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, c->stream);
    int rc = sab_saca_device(c, d_s, n, d_sa);
    cudaEventRecord(e1, c->stream);
    cudaStreamSynchronize(c->stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    sab_prof_collect(c);
    c->stats.total_ms = ms;
    if (rc != SAB_OK) {
        // a failed launch may have left the ticket counter out of step with its host mirror
        cudaMemsetAsync(c->d_ticket, 0, sizeof(u32) * 4, c->stream);
        cudaStreamSynchronize(c->stream);
        c->ticket_host = 0;
    }
    return rc;
}

static int sab_saca_multi(const u8* s, u64 n, u32* sa, int P, u32* bkt);  // sab_dist.cuh
static void sab_multi_shutdown();

// ------------------------------------------------------------------ extern "C"
extern "C" {

int32_t sab200_saca_device(const uint8_t* d_s, uint64_t n, uint32_t* d_sa, int32_t device) {
    if (!d_sa || (n > 0 && !d_s) || n > SAB200_MAX_LENGTH) {
        sab_set_error("sab200_saca_device: bad arguments (n=%llu)", (unsigned long long)n);
        return SAB_ERR_ARGS;
    }
    SabContext* c = sab_get_context(device);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_CUDA_TRY(cudaSetDevice(c->device));
    int rc = run_device(c, d_s, n, d_sa);
    g_last_stats = c->stats;
    return rc;
}

static int32_t saca_host(const uint8_t* s, uint64_t n, uint32_t* sa, uint32_t* bkt, int32_t ngpus);

int32_t sab200_saca(const uint8_t* s, uint64_t n, uint32_t* sa, int32_t ngpus) { return saca_host(s, n, sa, nullptr, ngpus); }

int32_t sab200_saca_buckets(const uint8_t* s, uint64_t n, uint32_t* sa, uint32_t* bkt, int32_t ngpus) {
    if (!bkt) {
        sab_set_error("sab200_saca_buckets: bkt is null");
        return SAB_ERR_ARGS;
    }
    return saca_host(s, n, sa, bkt, ngpus);
}

static int32_t saca_host(const uint8_t* s, uint64_t n, uint32_t* sa, uint32_t* bkt, int32_t ngpus) {
    if (ngpus == 0) ngpus = sab200_device_count() > SAB_MAX_RANKS ? SAB_MAX_RANKS : sab200_device_count();
    if (!sa || (n > 0 && !s) || n > SAB200_MAX_LENGTH || ngpus < 0 || ngpus > SAB_MAX_RANKS) {
        sab_set_error("sab200_saca: bad arguments (n=%llu, ngpus=%d)", (unsigned long long)n, (int)ngpus);
        return SAB_ERR_ARGS;
    }
    if (sab200_device_count() < 1) {
        sab_set_error("no CUDA device available; libsab200 has no CPU fallback");
        return SAB_ERR_CUDA;
    }
    if (ngpus > sab200_device_count()) {
        sab_set_error("sab200_saca: %d GPUs requested, %d visible", (int)ngpus, (int)sab200_device_count());
        return SAB_ERR_ARGS;
    }
    if (bkt && n < ((u64)1 << 20)) ngpus = 1;  // tiny texts may pack a single symbol per key: the table then needs the text
    if (ngpus > 1) return sab_saca_multi(s, n, sa, ngpus, bkt);
    SabContext* c = sab_get_context(0);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_CUDA_TRY(cudaSetDevice(c->device));
    // text and SA live at the top of the arena, the construction scratch below them
    const size_t text_bytes = sab_align_up((size_t)n + 64, 256);
    const size_t sa_bytes = sab_align_up(((size_t)n + 1) * sizeof(u32), 256);
    const size_t work = sab_saca_workspace_bytes(n);
    const size_t bkt_bytes = sab_align_up((size_t)SAB200_BKT_LEN * sizeof(u32), 256);
    SAB_TRY(sab_arena_reserve(c, work + text_bytes + sa_bytes + bkt_bytes + 512));
    u8* d_s = (u8*)(c->arena + sab_align_up(work, 256));
    u32* d_sa = (u32*)((char*)d_s + text_bytes);
    u32* d_bkt = (u32*)((char*)d_sa + sa_bytes);
    c->want_bkt = bkt ? d_bkt : nullptr;
    c->bkt_add_one = 1;
    double t0 = now_ms();
    SAB_TRY(sab_copy_h2d(c, d_s, s, n));
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    double t1 = now_ms();
    int rc = run_device(c, d_s, n, d_sa);
    c->want_bkt = nullptr;
    if (rc == SAB_OK) {
        double t2 = now_ms();
        if (bkt) {
            if (n == 0) {
                for (u32 i = 0; i < SAB200_BKT_LEN; ++i) bkt[i] = 1;  // src/sa.rs:98,112-116 on an empty text
            } else {
                SAB_CUDA_TRY(cudaMemcpyAsync(bkt, d_bkt, (size_t)SAB200_BKT_LEN * sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
            }
        }
        SAB_TRY(sab_copy_d2h(c, sa, d_sa, (n + 1) * sizeof(u32)));
        c->stats.h2d_ms = t1 - t0;
        c->stats.d2h_ms = now_ms() - t2;
    }
    g_last_stats = c->stats;
    return rc;
}

}  // extern "C"

// ---- buckets -------------------------------------------------------------------------------
static int buckets_device(SabContext* c, const u8* d_text, u64 n, u32* d_bkt) {
    cudaStream_t st = c->stream;
    SAB_CUDA_TRY(cudaMemsetAsync(d_bkt, 0, SAB_BKT_LEN * sizeof(u32), st));
    if (n > 0) {
        u32* d_hist = c->d_counters + 16;
        SAB_CUDA_TRY(cudaMemsetAsync(d_hist, 0, 256 * sizeof(u32), st));
        u64 blocks = div_up64(n, 256 * 64);
        const u64 bmax = (u64)c->sm_count * 8;
        if (blocks > bmax) blocks = bmax;
        SAB_LAUNCH(alphabet_hist_kernel, (unsigned)blocks, 256, 0, st, d_text, n, d_hist);
        SAB_LAUNCH_CHECK();
        SAB_CUDA_TRY(cudaMemcpyAsync(c->h_small + 64, d_hist, 256 * sizeof(u32), cudaMemcpyDeviceToHost, st));
        SAB_CUDA_TRY(cudaStreamSynchronize(st));
        u16 lut[256];
        u32 sigma = 0;
        for (int ch = 0; ch < 256; ++ch) {
            if (c->h_small[64 + ch]) ++sigma;
            lut[ch] = (u16)(c->h_small[64 + ch] ? sigma : 0);
        }
        u16* d_lut = (u16*)(c->d_counters + 16 + 256);
        memcpy(c->h_small + 384, lut, sizeof(lut));
        SAB_CUDA_TRY(cudaMemcpyAsync(d_lut, c->h_small + 384, sizeof(lut), cudaMemcpyHostToDevice, st));
        const size_t tab_bytes = (size_t)sigma * (sigma + 1) * sizeof(u32);
        const int dense = tab_bytes <= 160 * 1024;
        const size_t smem = dense ? tab_bytes : 0;
#ifndef SAB_EMU
        SAB_CUDA_TRY(cudaFuncSetAttribute(bucket_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
#endif
        u64 pblocks = div_up64(n, 256 * 64);
        const u64 pmax = (u64)c->sm_count * (smem > 64 * 1024 ? 1 : 4);
        if (pblocks > pmax) pblocks = pmax;
        SAB_LAUNCH(bucket_pairs_kernel, (unsigned)pblocks, 256, smem, st, d_text, n, (const u16*)d_lut, sigma, dense, d_bkt);
        SAB_LAUNCH_CHECK();
    }
    SAB_LAUNCH(bucket_scan_kernel, 1, 1024, 0, st, d_bkt);
    SAB_LAUNCH_CHECK();
    return SAB_OK;
}

extern "C" int32_t sab200_enable_buckets(const uint8_t* s, uint64_t n, uint32_t* bkt) {
    if (!bkt || (n > 0 && !s) || n > SAB200_MAX_LENGTH) {
        sab_set_error("sab200_enable_buckets: bad arguments");
        return SAB_ERR_ARGS;
    }
    SabContext* c = sab_get_context(0);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_CUDA_TRY(cudaSetDevice(c->device));
    const size_t text_bytes = sab_align_up((size_t)n + 64, 256);
    SAB_TRY(sab_arena_reserve(c, text_bytes + SAB_BKT_LEN * sizeof(u32) + 1024));
    u8* d_s = (u8*)c->arena;
    u32* d_bkt = (u32*)(c->arena + text_bytes);
    if (n) SAB_CUDA_TRY(cudaMemcpyAsync(d_s, s, n, cudaMemcpyHostToDevice, c->stream));
    SAB_TRY(buckets_device(c, d_s, n, d_bkt));
    SAB_CUDA_TRY(cudaMemcpyAsync(bkt, d_bkt, SAB_BKT_LEN * sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SAB_OK;
}

// ---- integrity check -----------------------------------------------------------------------
extern "C" int32_t sab200_check(const uint8_t* s, uint64_t n, const uint32_t* sa, uint64_t sa_len) {
    if (!sa || (n > 0 && !s) || n > SAB200_MAX_LENGTH) {
        sab_set_error("sab200_check: bad arguments");
        return SAB_ERR_ARGS;
    }
    if (sa_len != n + 1) return 0;  // src/sa.rs:73-75
    SabContext* c = sab_get_context(0);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_CUDA_TRY(cudaSetDevice(c->device));
    const size_t text_bytes = sab_align_up((size_t)n + 64, 256);
    const size_t arr_bytes = sab_align_up(((size_t)n + 2) * sizeof(u32), 256);
    SAB_TRY(sab_arena_reserve(c, text_bytes + 2 * arr_bytes + 1024));
    u8* d_s = (u8*)c->arena;
    u32* d_sa = (u32*)(c->arena + text_bytes);
    u32* d_isa = (u32*)(c->arena + text_bytes + arr_bytes);
    u32* d_flag = c->d_counters + 8;
    cudaStream_t st = c->stream;
    SAB_CUDA_TRY(cudaMemsetAsync(d_s, 0, text_bytes, st));
    if (n) SAB_CUDA_TRY(cudaMemcpyAsync(d_s, s, n, cudaMemcpyHostToDevice, st));
    SAB_CUDA_TRY(cudaMemcpyAsync(d_sa, sa, sa_len * sizeof(u32), cudaMemcpyHostToDevice, st));
    SAB_CUDA_TRY(cudaMemsetAsync(d_isa, 0xff, arr_bytes, st));
    SAB_CUDA_TRY(cudaMemsetAsync(d_flag, 0, sizeof(u32), st));
    const unsigned blocks = (unsigned)div_up64(sa_len, 256);
    SAB_LAUNCH(sufcheck_scatter_kernel, blocks, 256, 0, st, (const u32*)d_sa, sa_len, n, d_isa, d_flag);
    SAB_LAUNCH_CHECK();
    SAB_LAUNCH(sufcheck_order_kernel, blocks, 256, 0, st, (const u8*)d_s, (const u32*)d_sa, sa_len, n, (const u32*)d_isa, d_flag);
    SAB_LAUNCH_CHECK();
    SAB_CUDA_TRY(cudaMemcpyAsync(c->h_small, d_flag, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    return c->h_small[0] ? 0 : 1;
}

// ---- LCP array ------------------------------------------------------------------------------
extern "C" int32_t sab200_lcp_array(const uint8_t* s, uint64_t n, const uint32_t* sa, uint64_t sa_len, uint32_t* lcp) {
    if (!sa || !lcp || (n > 0 && !s) || n > SAB200_MAX_LENGTH || sa_len != n + 1) {
        sab_set_error("sab200_lcp_array: bad arguments (n=%llu, sa_len=%llu)", (unsigned long long)n, (unsigned long long)sa_len);
        return SAB_ERR_ARGS;
    }
    SabContext* c = sab_get_context(0);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_CUDA_TRY(cudaSetDevice(c->device));
    const size_t text_bytes = sab_align_up((size_t)n + 64, 256);
    const size_t arr_bytes = sab_align_up(((size_t)n + 2) * sizeof(u32), 256);
    SAB_TRY(sab_arena_reserve(c, text_bytes + 3 * arr_bytes + 1024));
    u8* d_s = (u8*)c->arena;
    u32* d_sa = (u32*)(c->arena + text_bytes);
    u32* d_isa = (u32*)(c->arena + text_bytes + arr_bytes);
    u32* d_lcp = (u32*)(c->arena + text_bytes + 2 * arr_bytes);
    cudaStream_t st = c->stream;
    if (n) SAB_CUDA_TRY(cudaMemcpyAsync(d_s, s, n, cudaMemcpyHostToDevice, st));
    SAB_CUDA_TRY(cudaMemcpyAsync(d_sa, sa, sa_len * sizeof(u32), cudaMemcpyHostToDevice, st));
    SAB_LAUNCH(lcp_isa_kernel, (unsigned)div_up64(sa_len, 256), 256, 0, st, (const u32*)d_sa, sa_len, d_isa);
    SAB_LAUNCH_CHECK();
    const u64 chunks = div_up64(n ? n : 1, SAB_LCP_CHUNK);
    SAB_LAUNCH(lcp_kasai_kernel, (unsigned)div_up64(chunks, 256), 256, 0, st, (const u8*)d_s, n, (const u32*)d_sa, (const u32*)d_isa, d_lcp);
    SAB_LAUNCH_CHECK();
    SAB_CUDA_TRY(cudaMemcpyAsync(lcp, d_lcp, sa_len * sizeof(u32), cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    return SAB_OK;
}

// ---- resident index + batched queries ------------------------------------------------------
struct SabReplica {
    SabContext* ctx = nullptr;
    u8* d_text = nullptr;
    u32* d_sa = nullptr;
    u32* d_bkt = nullptr;
    // prefix directory (sab_search.cuh PrefixDir): built here, never leaves the library
    u32* d_pdir = nullptr;
    u16* d_plut = nullptr;
    unsigned long long* d_probes = nullptr;
    PrefixDir pd;
    cudaStream_t stream2 = nullptr;  // second stream: the copies of one chunk of patterns run under the kernel of another
    // query scratch (grow-only)
    u8* d_pats = nullptr;
    size_t pats_cap = 0;
    u64* d_offs = nullptr;
    u32* d_out0 = nullptr;
    u32* d_out1 = nullptr;
    size_t np_cap = 0;
};
struct sab200_index {
    u64 n = 0;
    bool has_bkt = false;
    bool count_probes = false;
    std::vector<SabReplica> rep;
    std::mutex mu;
};

static void replica_free(SabReplica& r) {
    if (!r.ctx) return;
    cudaSetDevice(r.ctx->device);
    cudaFree(r.d_text);
    cudaFree(r.d_sa);
    cudaFree(r.d_bkt);
    cudaFree(r.d_pdir);
    cudaFree(r.d_plut);
    cudaFree(r.d_probes);
    cudaFree(r.d_pats);
    cudaFree(r.d_offs);
    cudaFree(r.d_out0);
    cudaFree(r.d_out1);
    if (r.stream2) cudaStreamDestroy(r.stream2);
    r = SabReplica();
}

// The prefix directory over the resident text and suffix array: alphabet from a byte histogram on the device, depth
// = the most symbols whose codes number at most n / 8 (between 2^16 and 2^27 entries), one sweep over the suffix
// array.  Running out of memory for it is not an error: the queries then bisect the whole bucket.
static int replica_build_prefix_dir(SabReplica& r, u64 n) {
    SabContext* c = r.ctx;
    cudaStream_t st = c->stream;
    std::lock_guard<std::mutex> lk(c->mu);
    u32* d_hist = c->d_counters + 16;
    SAB_CUDA_TRY(cudaMemsetAsync(d_hist, 0, 256 * sizeof(u32), st));
    u64 blocks = div_up64(n, 256 * 64);
    if (blocks > (u64)c->sm_count * 8) blocks = (u64)c->sm_count * 8;
    SAB_LAUNCH(alphabet_hist_kernel, (unsigned)blocks, 256, 0, st, (const u8*)r.d_text, n, d_hist);
    SAB_LAUNCH_CHECK();
    SAB_CUDA_TRY(cudaMemcpyAsync(c->h_small + 64, d_hist, 256 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    u16 lut[256];
    u32 sigma = 0;
    for (int ch = 0; ch < 256; ++ch) {
        const bool present = c->h_small[64 + ch] != 0;
        lut[ch] = (u16)(sigma | (present ? 0x8000u : 0u));
        if (present) ++sigma;
    }
    const u32 base = sigma < 2 ? 2 : sigma;
    u64 limit = n / 8;
    if (limit < (1ull << 16)) limit = 1ull << 16;
    if (limit > (1ull << 27)) limit = 1ull << 27;
    u32 depth = 1;
    u64 entries = base;
    while (depth < SAB_PDIR_MAXD && entries * base <= limit) {
        entries *= base;
        ++depth;
    }
    if (cudaMalloc(&r.d_pdir, (entries + 1) * sizeof(u32)) != cudaSuccess) {
        cudaGetLastError();
        r.d_pdir = nullptr;
        return SAB_OK;
    }
    SAB_CUDA_TRY(cudaMalloc(&r.d_plut, sizeof(lut)));
    memcpy(c->h_small + 384, lut, sizeof(lut));  // pinned staging
    SAB_CUDA_TRY(cudaMemcpyAsync(r.d_plut, c->h_small + 384, sizeof(lut), cudaMemcpyHostToDevice, st));
    u64 fblocks = div_up64(entries + 1, 256 * 8);
    if (fblocks > (u64)c->sm_count * 16) fblocks = (u64)c->sm_count * 16;
    SAB_LAUNCH(fill_u32_kernel, (unsigned)fblocks, 256, 0, st, r.d_pdir, entries + 1, (u32)(n + 1));
    SAB_LAUNCH_CHECK();
    SAB_LAUNCH(prefix_dir_kernel, (unsigned)div_up64(n + 1, 256), 256, 0, st, (const u8*)r.d_text, n, (const u32*)r.d_sa, n + 1,
               (const u16*)r.d_plut, base, depth, r.d_pdir);
    SAB_LAUNCH_CHECK();
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    r.pd.dir = r.d_pdir;
    r.pd.lut = r.d_plut;
    r.pd.sigma = base;
    r.pd.depth = depth;
    r.pd.pw[0] = 1;
    for (u32 t = 1; t <= SAB_PDIR_MAXD; ++t) r.pd.pw[t] = t <= depth ? r.pd.pw[t - 1] * base : 0;
    return SAB_OK;
}

static int replica_init(SabReplica& r, int device, const u8* s, u64 n, const u32* sa, const u32* bkt) {
    r.ctx = sab_get_context(device);
    if (!r.ctx) return SAB_ERR_CUDA;
    SAB_CUDA_TRY(cudaSetDevice(device));
    cudaStream_t st = r.ctx->stream;
    const size_t text_bytes = sab_align_up((size_t)n + 64, 256);
    SAB_CUDA_TRY(cudaMalloc(&r.d_text, text_bytes));
    SAB_CUDA_TRY(cudaMalloc(&r.d_sa, ((size_t)n + 1) * sizeof(u32)));
    SAB_CUDA_TRY(cudaMemsetAsync(r.d_text, 0, text_bytes, st));
    if (n) SAB_CUDA_TRY(cudaMemcpyAsync(r.d_text, s, n, cudaMemcpyHostToDevice, st));
    SAB_CUDA_TRY(cudaMemcpyAsync(r.d_sa, sa, ((size_t)n + 1) * sizeof(u32), cudaMemcpyHostToDevice, st));
    if (bkt) {
        SAB_CUDA_TRY(cudaMalloc(&r.d_bkt, SAB_BKT_LEN * sizeof(u32)));
        SAB_CUDA_TRY(cudaMemcpyAsync(r.d_bkt, bkt, SAB_BKT_LEN * sizeof(u32), cudaMemcpyHostToDevice, st));
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    memset(&r.pd, 0, sizeof(r.pd));
    SAB_CUDA_TRY(cudaStreamCreateWithFlags(&r.stream2, cudaStreamNonBlocking));
    SAB_CUDA_TRY(cudaMalloc(&r.d_probes, sizeof(unsigned long long)));
    SAB_CUDA_TRY(cudaMemsetAsync(r.d_probes, 0, sizeof(unsigned long long), st));
    const char* off = getenv("SAB_SEARCH_DIR");
    if (n > 0 && !(off && off[0] == '0')) SAB_TRY(replica_build_prefix_dir(r, n));
    return SAB_OK;
}

extern "C" sab200_index* sab200_index_create(const uint8_t* s, uint64_t n, const uint32_t* sa, uint64_t sa_len,
                                             const uint32_t* bkt_or_null, int32_t ngpus) {
    if (!sa || (n > 0 && !s) || n > SAB200_MAX_LENGTH || ngpus < 1 || ngpus > SAB_MAX_DEVICES) {
        sab_set_error("sab200_index_create: bad arguments");
        return nullptr;
    }
    if (sa_len != n + 1) {
        sab_set_error("sab200_index_create: the suffix array has %llu entries, the text needs %llu",
                      (unsigned long long)sa_len, (unsigned long long)(n + 1));
        return nullptr;
    }
    if (ngpus > sab200_device_count()) {
        sab_set_error("sab200_index_create: %d GPUs requested, %d visible", (int)ngpus, (int)sab200_device_count());
        return nullptr;
    }
    sab200_index* ix = new (std::nothrow) sab200_index();
    if (!ix) return nullptr;
    ix->n = n;
    ix->has_bkt = bkt_or_null != nullptr;
    ix->rep.resize(ngpus);
    for (int d = 0; d < ngpus; ++d) {
        if (replica_init(ix->rep[d], d, s, n, sa, bkt_or_null) != SAB_OK) {
            for (auto& r : ix->rep) replica_free(r);
            delete ix;
            return nullptr;
        }
    }
    return ix;
}

extern "C" void sab200_index_destroy(sab200_index* ix) {
    if (!ix) return;
    for (auto& r : ix->rep) replica_free(r);
    delete ix;
}

template <int MODE>
static int launch_search(SabReplica& r, u64 n, const u8* d_pats_adj, const u64* d_offs, u64 np, u32* d_out0, u32* d_out1,
                         bool count_probes = false, cudaStream_t st = nullptr) {
    if (!st) st = r.ctx->stream;
    SearchArgs a;
    a.pd = r.pd;
    a.probes = count_probes ? r.d_probes : nullptr;
    a.text = r.d_text;
    a.n = n;
    a.sa = r.d_sa;
    a.bkt = r.d_bkt;
    a.pats = d_pats_adj;
    a.offs = d_offs;
    a.np = np;
    a.out0 = d_out0;
    a.out1 = d_out1;
    const u64 threads = np * SAB_SEARCH_G;
    SAB_LAUNCH((search_kernel<SAB_SEARCH_G, MODE>), (unsigned)div_up64(threads, SAB_SEARCH_THREADS), SAB_SEARCH_THREADS, 0, st, a);
    SAB_LAUNCH_CHECK();
    return SAB_OK;
}

#ifndef SAB_SEARCH_CHUNK
#define SAB_SEARCH_CHUNK (1ull << 20)  // patterns per pipelined chunk
#endif
static int replica_reserve(SabReplica& r, size_t pat_bytes, size_t np) {
    if (r.pats_cap < pat_bytes + 64) {
        cudaFree(r.d_pats);
        r.d_pats = nullptr;
        r.pats_cap = 0;
        const size_t want = sab_align_up(pat_bytes + pat_bytes / 4 + 4096, 256);
        SAB_CUDA_TRY(cudaMalloc(&r.d_pats, want));
        SAB_CUDA_TRY(cudaMemset(r.d_pats, 0, want));
        r.pats_cap = want;
    }
    if (r.np_cap < np + np / SAB_SEARCH_CHUNK + 2) {  // every pipelined chunk keeps its own end offset
        cudaFree(r.d_offs);
        cudaFree(r.d_out0);
        cudaFree(r.d_out1);
        r.d_offs = nullptr;
        r.d_out0 = r.d_out1 = nullptr;
        r.np_cap = 0;
        const size_t want = np + np / 4 + 1024;
        SAB_CUDA_TRY(cudaMalloc(&r.d_offs, want * sizeof(u64)));
        SAB_CUDA_TRY(cudaMalloc(&r.d_out0, want * sizeof(u32)));
        SAB_CUDA_TRY(cudaMalloc(&r.d_out1, want * sizeof(u32)));
        r.np_cap = want;
    }
    return SAB_OK;
}

// mode 0 search_all (out0 = lo, out1 = hi), 1 contains (out0 = u8 flags), 2 search_lcp (start, end)
static int search_batch(sab200_index* ix, int mode, const u8* pats, const u64* offs, u64 np, void* out0, void* out1) {
    if (!ix || (np > 0 && (!offs || !out0)) || (mode != 1 && np > 0 && !out1)) {
        sab_set_error("batched search: bad arguments");
        return SAB_ERR_ARGS;
    }
    if (np == 0) return SAB_OK;
    if (np > 0x7fffffffull) {
        sab_set_error("batched search: at most 2^31-1 patterns per call");
        return SAB_ERR_ARGS;
    }
    std::lock_guard<std::mutex> lk(ix->mu);
    const int P = (int)ix->rep.size();
    const u64 per = div_up64(np, (u64)P);
    // enqueue on every replica, then drain
    for (int d = 0; d < P; ++d) {
        const u64 q0 = (u64)d * per, q1 = (q0 + per < np) ? q0 + per : np;
        if (q0 >= q1) continue;
        SabReplica& r = ix->rep[d];
        SAB_CUDA_TRY(cudaSetDevice(r.ctx->device));
        const u64 b0 = offs[q0], b1 = offs[q1];
        const u64 cnt = q1 - q0;
        SAB_TRY(replica_reserve(r, (size_t)(b1 - b0), (size_t)cnt));
        // Chunks of SAB_SEARCH_CHUNK patterns alternate between two streams: the pattern upload of one chunk and
        // the result download of another run under the kernel of a third (pinned caller buffers; PCIe is full
        // duplex).  Every chunk owns its sub-ranges of the scratch buffers, so the streams never meet.
        const u8* adj = r.d_pats - b0;  // the kernel indexes patterns by their absolute offsets
        int which = 0;
        for (u64 c0 = 0; c0 < cnt; c0 += SAB_SEARCH_CHUNK, which ^= 1) {
            const u64 c1 = c0 + SAB_SEARCH_CHUNK < cnt ? c0 + SAB_SEARCH_CHUNK : cnt;
            const u64 cc = c1 - c0;
            const u64 cb0 = offs[q0 + c0], cb1 = offs[q0 + c1];
            cudaStream_t st = (which && r.stream2) ? r.stream2 : r.ctx->stream;
            if (cb1 > cb0) SAB_CUDA_TRY(cudaMemcpyAsync(r.d_pats + (cb0 - b0), pats + cb0, cb1 - cb0, cudaMemcpyHostToDevice, st));
            // chunk c reads offs[c0 .. c1]; entry c1 is also the first of the next chunk, which lives one slot further
            u64* d_offs = r.d_offs + c0 + c0 / SAB_SEARCH_CHUNK;
            SAB_CUDA_TRY(cudaMemcpyAsync(d_offs, offs + q0 + c0, (cc + 1) * sizeof(u64), cudaMemcpyHostToDevice, st));
            u32* o0 = r.d_out0 + c0;
            u32* o1 = r.d_out1 + c0;
            if (mode == 0) SAB_TRY(launch_search<0>(r, ix->n, adj, d_offs, cc, o0, o1, ix->count_probes, st));
            else if (mode == 1) SAB_TRY(launch_search<1>(r, ix->n, adj, d_offs, cc, (u32*)((u8*)r.d_out0 + c0), o1, false, st));
            else SAB_TRY(launch_search<2>(r, ix->n, adj, d_offs, cc, o0, o1, false, st));
            if (mode == 1) {
                SAB_CUDA_TRY(cudaMemcpyAsync((u8*)out0 + q0 + c0, (u8*)r.d_out0 + c0, cc, cudaMemcpyDeviceToHost, st));
            } else {
                SAB_CUDA_TRY(cudaMemcpyAsync((u32*)out0 + q0 + c0, o0, cc * sizeof(u32), cudaMemcpyDeviceToHost, st));
                SAB_CUDA_TRY(cudaMemcpyAsync((u32*)out1 + q0 + c0, o1, cc * sizeof(u32), cudaMemcpyDeviceToHost, st));
            }
        }
    }
    for (int d = 0; d < P; ++d) {
        SabReplica& r = ix->rep[d];
        SAB_CUDA_TRY(cudaSetDevice(r.ctx->device));
        SAB_CUDA_TRY(cudaStreamSynchronize(r.ctx->stream));
        if (r.stream2) SAB_CUDA_TRY(cudaStreamSynchronize(r.stream2));
    }
    return SAB_OK;
}

extern "C" int32_t sab200_search_all_batch(sab200_index* ix, const uint8_t* pats, const uint64_t* offs, uint64_t np, uint32_t* lo,
                                uint32_t* hi) {
    return search_batch(ix, 0, pats, offs, np, lo, hi);
}
extern "C" int32_t sab200_contains_batch(sab200_index* ix, const uint8_t* pats, const uint64_t* offs, uint64_t np, uint8_t* out) {
    return search_batch(ix, 1, pats, offs, np, out, nullptr);
}
extern "C" int32_t sab200_search_lcp_batch(sab200_index* ix, const uint8_t* pats, const uint64_t* offs, uint64_t np,
                                uint32_t* start, uint32_t* end) {
    return search_batch(ix, 2, pats, offs, np, start, end);
}

// Device-resident variant for timing the kernel alone: d_pats / d_offs / d_lo / d_hi live on the
// device of replica 0; d_pats must be followed by >= 8 readable bytes.
extern "C" int32_t sab200_search_all_batch_device(sab200_index* ix, const uint8_t* d_pats, const uint64_t* d_offs, uint64_t np,
                                       uint32_t* d_lo, uint32_t* d_hi) {
    if (!ix || !d_offs || !d_lo || !d_hi) return SAB_ERR_ARGS;
    if (np == 0) return SAB_OK;
    std::lock_guard<std::mutex> lk(ix->mu);
    SabReplica& r = ix->rep[0];
    SAB_CUDA_TRY(cudaSetDevice(r.ctx->device));
    SAB_TRY(launch_search<0>(r, ix->n, d_pats, d_offs, np, d_lo, d_hi, ix->count_probes));
    SAB_CUDA_TRY(cudaStreamSynchronize(r.ctx->stream));
    return SAB_OK;
}

// Probe counting for the roofline of the search kernel (bench.py): while on, search_all adds the suffix comparisons
// of every pattern to a device counter (one atomic per pattern -- not for timed runs); returns the sum over the
// replicas since the index was created.
extern "C" uint64_t sab200_index_probes(sab200_index* ix, int32_t count_on) {
    if (!ix) return 0;
    std::lock_guard<std::mutex> lk(ix->mu);
    ix->count_probes = count_on != 0;
    u64 total = 0;
    for (auto& r : ix->rep) {
        if (!r.d_probes) continue;
        unsigned long long v = 0;
        cudaSetDevice(r.ctx->device);
        cudaStreamSynchronize(r.ctx->stream);
        if (cudaMemcpy(&v, r.d_probes, sizeof(v), cudaMemcpyDeviceToHost) == cudaSuccess) total += v;
    }
    return total;
}

// Layout of the prefix directory of replica 0: *sigma = base, *depth = symbols per code; returns the number of
// entries (0: no directory).
extern "C" uint64_t sab200_index_directory(sab200_index* ix, uint32_t* sigma, uint32_t* depth) {
    if (!ix || ix->rep.empty() || !ix->rep[0].pd.dir) return 0;
    if (sigma) *sigma = ix->rep[0].pd.sigma;
    if (depth) *depth = ix->rep[0].pd.depth;
    return ix->rep[0].pd.pw[ix->rep[0].pd.depth];
}

// ---- pack serialisation ---------------------------------------------------------------------
extern "C" uint64_t sab200_pack_bound(uint64_t sa_len) { return 16 + ((sa_len + 127) / 128) * 512; }

extern "C" int32_t sab200_pack(const uint32_t* sa, uint64_t sa_len, uint8_t* out, uint64_t out_cap, uint64_t* out_len) {
    if (!sa || !out || !out_len || sa_len == 0 || sa_len > 0xffffffffull) {  // src/packed_sa.rs:18
        sab_set_error("sab200_pack: bad arguments");
        return SAB_ERR_ARGS;
    }
    const unsigned bits = sab_pack_bits(sa_len);
    const u64 blocks = (sa_len + 127) / 128;
    const u64 words = blocks * 4ull * bits;
    if (out_cap < 16 + words * 4) {
        sab_set_error("sab200_pack: output buffer too small (need %llu bytes)", (unsigned long long)(16 + words * 4));
        return SAB_ERR_ARGS;
    }
    SabContext* c = sab_get_context(0);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_CUDA_TRY(cudaSetDevice(c->device));
    const size_t in_bytes = sab_align_up((size_t)sa_len * 4 + 64, 256);
    SAB_TRY(sab_arena_reserve(c, in_bytes + (size_t)words * 4 + 1024));
    u32* d_sa = (u32*)c->arena;
    u32* d_out = (u32*)(c->arena + in_bytes);
    cudaStream_t st = c->stream;
    SAB_CUDA_TRY(cudaMemcpyAsync(d_sa, sa, sa_len * 4, cudaMemcpyHostToDevice, st));
    if (words) {
        SAB_LAUNCH(pack_blocks_kernel, (unsigned)div_up64(words, 256), 256, 0, st, (const u32*)d_sa, sa_len, bits, words, d_out);
        SAB_LAUNCH_CHECK();
        SAB_CUDA_TRY(cudaMemcpyAsync(out + 16, d_out, words * 4, cudaMemcpyDeviceToHost, st));
    }
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    u64 data_len = words * 4;
    if (sa_len % 128 != 0) {  // src/packed_sa.rs:41-45: trailing zero bytes of the padded last block are dropped
        const u64 chunk = 16ull * bits, start = data_len - chunk;
        u64 t = chunk;
        while (t > 0 && out[16 + start + t - 1] == 0) --t;
        data_len = start + t;
    }
    const u32 magic = SAB_PACK_MAGIC, length = (u32)sa_len;
    memcpy(out, &magic, 4);
    memcpy(out + 4, &length, 4);
    memcpy(out + 8, &data_len, 8);
    *out_len = 16 + data_len;
    return SAB_OK;
}

extern "C" int32_t sab200_unpack(const uint8_t* bytes, uint64_t nbytes, uint32_t* sa, uint64_t sa_cap, uint64_t* sa_len) {
    if (!bytes || !sa || !sa_len || nbytes < 16) {
        sab_set_error("sab200_unpack: bad arguments");
        return SAB_ERR_ARGS;
    }
    u32 magic, length;
    u64 dlen;
    memcpy(&magic, bytes, 4);
    memcpy(&length, bytes + 4, 4);
    memcpy(&dlen, bytes + 8, 8);
    const unsigned bits = sab_pack_bits(length);
    const u64 blocks = ((u64)length + 127) / 128, words = blocks * 4ull * bits;
    if (magic != SAB_PACK_MAGIC || dlen != nbytes - 16 || dlen > words * 4 || length == 0 || length > sa_cap) {
        sab_set_error("sab200_unpack: not a packed suffix array (magic %08x, length %u, %llu data bytes)", magic, length,
                      (unsigned long long)dlen);
        return SAB_ERR_ARGS;
    }
    SabContext* c = sab_get_context(0);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_CUDA_TRY(cudaSetDevice(c->device));
    const size_t in_bytes = sab_align_up((size_t)words * 4 + 64, 256);
    SAB_TRY(sab_arena_reserve(c, in_bytes + (size_t)length * 4 + 1024));
    u32* d_in = (u32*)c->arena;
    u32* d_sa = (u32*)(c->arena + in_bytes);
    cudaStream_t st = c->stream;
    SAB_CUDA_TRY(cudaMemsetAsync(d_in, 0, in_bytes, st));  // re-pads the trimmed tail (src/packed_sa.rs:80-85)
    if (dlen) SAB_CUDA_TRY(cudaMemcpyAsync(d_in, bytes + 16, dlen, cudaMemcpyHostToDevice, st));
    SAB_LAUNCH(unpack_blocks_kernel, (unsigned)div_up64(length, 256), 256, 0, st, (const u32*)d_in, (u64)length, bits, d_sa);
    SAB_LAUNCH_CHECK();
    SAB_CUDA_TRY(cudaMemcpyAsync(sa, d_sa, (size_t)length * 4, cudaMemcpyDeviceToHost, st));
    SAB_CUDA_TRY(cudaStreamSynchronize(st));
    *sa_len = length;
    return SAB_OK;
}

extern "C" {

int32_t sab200_get_stats(sab200_stats* out) {
    if (!out) return SAB_ERR_ARGS;
    memcpy(out, &g_last_stats, sizeof(*out));
    return SAB_OK;
}

void sab200_set_profiling(int32_t on) { g_profiling = on != 0; }

const char* sab200_last_error(void) { return g_err; }

int32_t sab200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* sab200_version(void) {
#ifdef SAB_EMU
    return "sab200 0.1.0 (SIMT emulator build -- tests only)";
#else
    return "sab200 0.1.0 (sm_100a)";
#endif
}

void sab200_shutdown(void) {
    sab_multi_shutdown();
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (int d = 0; d < SAB_MAX_DEVICES; ++d) {
        if (g_ctx[d]) {
            sab_context_destroy(g_ctx[d]);
            g_ctx[d] = nullptr;
        }
    }
}

}  // extern "C"

#include "sab_dist.cuh"
