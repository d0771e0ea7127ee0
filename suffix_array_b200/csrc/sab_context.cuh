// sab_context.cuh -- per-device context: stream, scratch for the radix sort and the chained scans,
// and a grow-only arena for the big arrays.  Device memory, streams and events are the only
// resources the library owns (SURVEY.md 8b "Ownership"); no host pointer is retained.
#pragma once
#include <stdarg.h>

#include <mutex>
#include <vector>

#include "sab_common.cuh"
#include "sab_scan.cuh"

#define SAB_MAX_ROUNDS 64

// mirrors sab200_stats in include/sab200.h (kept in sync by a static_assert in sab_api.cu)
struct SabStats {
    u64 n;                  // text length of the last construction
    u32 sigma;              // distinct byte values
    u32 bits_per_symbol;    // b
    u32 symbols_per_key;    // k
    u32 rounds;             // doubling rounds executed (0 = settled by the initial sort)
    u64 active[SAB_MAX_ROUNDS];   // active suffixes entering round r (r = 0: after the initial sort)
    u32 passes[SAB_MAX_ROUNDS];   // radix passes executed in the sort of round r (index 0 = initial)
    u64 radix_pass_launches;      // onesweep launches
    u64 radix_pass_records;       // sum over launches of records moved
    u64 radix_pass_bytes;         // sum over launches of 2*(K+V)*m  (algorithmic bytes)
    double radix_pass_ms;         // sum of CUDA-event durations of those launches (profiling on)
    double hist_ms;               // radix histogram kernels
    double pack_ms;               // alphabet + key packing
    double rank_ms;               // head-flag / rank / compaction kernels
    double gather_ms;             // rank[i+h] gathers
    double total_ms;              // whole device pipeline (events around it)
    double h2d_ms, d2h_ms;        // host entry only
    u64 kernel_launches;          // all kernels launched by the last call
    double group_sort_ms;         // in-group sorts of the rounds (group_sort_kernel + scatter-back)
    u64 group_sort_records;       // records that went through group_sort_kernel
    u64 group_big_records;        // ... of which in groups too large for it (radix-sorted)
};

struct SabEventPair {
    cudaEvent_t a, b;
    int kind;  // 0 radix pass, 1 hist, 2 pack, 3 rank, 4 gather, 5 group sort
};

struct SabContext {
    int device = 0;
    bool ready = false;
    cudaStream_t stream = 0;
    int sm_count = 148;
    // radix scratch
    u64* d_ghist = nullptr;   // [8][256]
    u64* d_gbase = nullptr;   // [8][256]
    u32* d_skip = nullptr;    // [8]
    u32* h_small = nullptr;   // pinned, 4096 words: [0,16) counters read back, [32] sentinel staging, [64,320) byte
                              // histogram, [384,512) code table staging, [1024,2048) 512 u64 of count / base staging
    u64* d_lookback = nullptr;
    size_t lookback_tiles = 0;
    u32* d_ticket = nullptr;
    u32 ticket_host = 0;
    u32 lb_epoch = 0;
    // chained-scan scratch (3 x u32 states at most)
    ScanSlot* d_scan_slots = nullptr;  // chained-scan status, one 16-byte slot per tile
    size_t scan_tiles = 0;
    u32 scan_epoch = 0;
    u32* d_counters = nullptr;  // 1024 words: [0,16) device scalars (counts of the scans), [16,272) byte histogram,
                                // [272,400) code table (256 x u16)
    // arena
    char* arena = nullptr;
    size_t arena_bytes = 0;
    size_t arena_used = 0;
    size_t arena_want_seen = 0;
    void* bounce[2] = {nullptr, nullptr};        // pinned bounce buffers for pageable caller memory (sab_copy_*)
    cudaEvent_t bounce_ev[2] = {nullptr, nullptr};
    u32* want_bkt = nullptr;     // device buffer for the fused bucket table of the running construction (or null)
    u32 bkt_add_one = 1;  // request the arena was last sized for (it may have been capped by the free memory)
    // profiling
    bool profiling = false;
    std::vector<SabEventPair> events;
    std::vector<cudaEvent_t> event_pool;
    SabStats stats;
    std::mutex mu;
};

SabContext* sab_get_context(int device);
int sab_context_init(SabContext* c);
void sab_context_destroy(SabContext* c);
int sab_arena_reserve(SabContext* c, size_t bytes);
int sab_ensure_lookback(SabContext* c, size_t tiles);
int sab_ensure_scan(SabContext* c, size_t tiles);

static inline size_t sab_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocation inside the arena (call sab_arena_reserve first)
template <typename T>
static inline T* sab_arena_take(SabContext* c, size_t count) {
    size_t off = sab_align_up(c->arena_used, 256);
    c->arena_used = off + count * sizeof(T);
    return (T*)(c->arena + off);
}

cudaEvent_t sab_event_get(SabContext* c);
int sab_copy_h2d(SabContext* c, void* d_dst, const void* h_src, size_t bytes);
int sab_copy_d2h(SabContext* c, void* h_dst, const void* d_src, size_t bytes);
void sab_prof_begin(SabContext* c, int kind);
void sab_prof_end(SabContext* c);
void sab_prof_collect(SabContext* c);

template <typename T>
static inline TileState<T> sab_tile_state(SabContext* c, size_t tiles) {
    (void)tiles;
    TileState<T> st;
    st.slots = c->d_scan_slots;
    if (++c->scan_epoch >= (1u << 30)) {  // wrap: stale tags of 2^30 launches ago must not match
        cudaMemsetAsync(c->d_scan_slots, 0, c->scan_tiles * sizeof(ScanSlot), c->stream);
        c->scan_epoch = 1;
    }
    st.epoch = c->scan_epoch;
    return st;
}
