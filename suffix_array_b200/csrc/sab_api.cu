// sab_api.cu -- context management and the extern "C" surface declared in include/sab200.h.
// Single translation unit: the kernels live in the .cuh headers included below.
#include "../../include/sab200.h"
#include "sab_common.cuh"
#include "sab_context.cuh"
#include "sab_saca.cuh"

#include <chrono>
#include <cstdlib>
#include <new>

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
    SAB_CUDA_TRY(cudaMallocHost(&c->h_small, sizeof(u32) * 1024));
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
    cudaFree(c->d_lookback);
    cudaFree(c->d_ticket);
    cudaFree(c->d_scan_flags);
    cudaFree(c->d_scan_partial);
    cudaFree(c->d_scan_inclusive);
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

int sab_ensure_lookback(SabContext* c, size_t tiles) {
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
    if (c->d_scan_flags) {
        SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_scan_flags);
        cudaFree(c->d_scan_partial);
        cudaFree(c->d_scan_inclusive);
        c->d_scan_flags = c->d_scan_partial = c->d_scan_inclusive = nullptr;
        c->scan_tiles = 0;
    }
    const size_t want = tiles + tiles / 8 + 64;
    SAB_CUDA_TRY(cudaMalloc(&c->d_scan_flags, want * sizeof(u32)));
    SAB_CUDA_TRY(cudaMalloc(&c->d_scan_partial, want * 16));
    SAB_CUDA_TRY(cudaMalloc(&c->d_scan_inclusive, want * 16));
    SAB_CUDA_TRY(cudaMemsetAsync(c->d_scan_flags, 0, want * sizeof(u32), c->stream));
    c->scan_tiles = want;
    c->scan_epoch = 0;
    return SAB_OK;
}

// ------------------------------------------------------------------ per-launch event timing
static cudaEvent_t sab_event_get(SabContext* c) {
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
        }
        c->event_pool.push_back(p.a);
        c->event_pool.push_back(p.b);
    }
    c->events.clear();
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

int32_t sab200_saca(const uint8_t* s, uint64_t n, uint32_t* sa, int32_t ngpus) {
    if (!sa || (n > 0 && !s) || n > SAB200_MAX_LENGTH || ngpus != 1) {
        sab_set_error("sab200_saca: bad arguments (n=%llu, ngpus=%d)", (unsigned long long)n, (int)ngpus);
        return SAB_ERR_ARGS;
    }
    SabContext* c = sab_get_context(0);
    if (!c) return SAB_ERR_CUDA;
    std::lock_guard<std::mutex> lk(c->mu);
    SAB_CUDA_TRY(cudaSetDevice(c->device));
    // text and SA live at the top of the arena, the construction scratch below them
    const size_t text_bytes = sab_align_up((size_t)n + 64, 256);
    const size_t sa_bytes = sab_align_up(((size_t)n + 1) * sizeof(u32), 256);
    const size_t work = sab_saca_workspace_bytes(n);
    SAB_TRY(sab_arena_reserve(c, work + text_bytes + sa_bytes + 512));
    u8* d_s = (u8*)(c->arena + sab_align_up(work, 256));
    u32* d_sa = (u32*)((char*)d_s + text_bytes);
    double t0 = now_ms();
    if (n) SAB_CUDA_TRY(cudaMemcpyAsync(d_s, s, n, cudaMemcpyHostToDevice, c->stream));
    SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
    double t1 = now_ms();
    int rc = run_device(c, d_s, n, d_sa);
    if (rc == SAB_OK) {
        double t2 = now_ms();
        SAB_CUDA_TRY(cudaMemcpyAsync(sa, d_sa, (n + 1) * sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
        SAB_CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->stats.h2d_ms = t1 - t0;
        c->stats.d2h_ms = now_ms() - t2;
    }
    g_last_stats = c->stats;
    return rc;
}

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
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (int d = 0; d < SAB_MAX_DEVICES; ++d) {
        if (g_ctx[d]) {
            sab_context_destroy(g_ctx[d]);
            g_ctx[d] = nullptr;
        }
    }
}

}  // extern "C"
