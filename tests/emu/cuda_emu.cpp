// cuda_emu.cpp -- fiber scheduler behind tests/emu/cuda_emu.h (TEST INFRASTRUCTURE ONLY).
#include "cuda_emu.h"

#include <sys/mman.h>
#include <time.h>

#include <vector>

emu_uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

double emu_now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

extern "C" void emu_swap(void** save_sp, void* new_sp);
asm(R"(
.text
.globl emu_swap
.type emu_swap,@function
emu_swap:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_swap,.-emu_swap
)");

namespace emu {

enum { RUNNABLE = 0, AT_BLOCK = 1, AT_WARP = 2, DONE = 3 };
static const size_t STACK_BYTES = 48 * 1024;
static const int MAX_THREADS = 1024;

struct Fiber {
    void* sp;
    int state;
    int tid;
    unsigned wait_mask;  // lanes this fiber rendezvouses with (AT_WARP)
};
struct WarpState {
    uint64_t slots[32];
    unsigned arrived;  // bit per lane waiting at a warp rendezvous
    unsigned live;
};
struct SharedVar {
    const void* key;
    void* ptr;
};
struct Block {
    int bid;
    int nthreads;
    int live;
    int at_block;
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    std::vector<SharedVar> shared;
    void* dyn;
    char* stacks;  // slot in the stack pool
    int slot;
};

static int g_window = 3;
static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static void* g_sched_sp;
static Block* g_block;
static Fiber* g_fiber;
static const std::function<void()>* g_body;
static char* g_pool = nullptr;
static int g_pool_slots = 0;
static size_t g_dyn_bytes;

void set_seed(uint64_t s) { g_rng = s * 0x9E3779B97F4A7C15ull + 0x1234567ull; }
void set_window(int w) { g_window = w < 1 ? 1 : (w > 8 ? 8 : w); }
static uint64_t rnd() {
    g_rng += 0x9E3779B97F4A7C15ull;
    uint64_t z = g_rng;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static void to_scheduler() { emu_swap(&g_fiber->sp, g_sched_sp); }

static void fiber_main() {
    (*g_body)();
    Block* b = g_block;
    Fiber* f = g_fiber;
    f->state = DONE;
    b->live--;
    b->warps[f->tid >> 5].live &= ~(1u << (f->tid & 31));
    to_scheduler();
    fprintf(stderr, "emu: resumed a finished fiber\n");
    abort();
}

void yield_spin() { to_scheduler(); }
void sync_block() {
    g_fiber->state = AT_BLOCK;
    g_block->at_block++;
    to_scheduler();
}
void sync_warp(unsigned mask) {
    g_fiber->state = AT_WARP;
    g_fiber->wait_mask = mask;
    g_block->warps[g_fiber->tid >> 5].arrived |= 1u << (g_fiber->tid & 31);
    to_scheduler();
}
// release every group of lanes whose members (still alive) have all arrived with the same mask
static bool release_warp(Block& b, int w) {
    WarpState& ws = b.warps[w];
    bool any = false;
    unsigned pending = ws.arrived;
    while (pending) {
        const int lane = __builtin_ctz(pending);
        Fiber& f = b.fibers[w * 32 + lane];
        const unsigned need = f.wait_mask & ws.live;
        pending &= ~need;
        pending &= ~(1u << lane);
        if ((ws.arrived & need) != need) continue;
        bool same = true;
        for (unsigned m = need; m; m &= m - 1)
            if (b.fibers[w * 32 + __builtin_ctz(m)].wait_mask != f.wait_mask) same = false;
        if (!same) {
            fprintf(stderr, "emu: lanes of one warp collective disagree on the member mask\n");
            abort();
        }
        for (unsigned m = need; m; m &= m - 1) b.fibers[w * 32 + __builtin_ctz(m)].state = RUNNABLE;
        ws.arrived &= ~need;
        any = true;
    }
    return any;
}
uint64_t* warp_slots() { return g_block->warps[g_fiber->tid >> 5].slots; }
uint32_t warp_live_mask() { return g_block->warps[g_fiber->tid >> 5].live; }
void* dyn_smem() { return g_block->dyn; }
void* shared_get(const void* key, size_t bytes) {
    for (auto& v : g_block->shared)
        if (v.key == key) return v.ptr;
    void* p = aligned_alloc(16, (bytes + 15) & ~(size_t)15);
    memset(p, 0xA5, bytes);  // shared memory is not initialised on a GPU either
    g_block->shared.push_back({key, p});
    return p;
}

static void start_block(Block& b, int bid, int nthreads, int slot) {
    b.bid = bid;
    b.nthreads = nthreads;
    b.live = nthreads;
    b.at_block = 0;
    b.slot = slot;
    b.stacks = g_pool + (size_t)slot * MAX_THREADS * STACK_BYTES;
    b.fibers.assign(nthreads, Fiber());
    b.warps.assign((nthreads + 31) / 32, WarpState());
    b.shared.clear();
    b.dyn = g_dyn_bytes ? aligned_alloc(128, (g_dyn_bytes + 127) & ~(size_t)127) : nullptr;
    if (b.dyn) memset(b.dyn, 0xA5, g_dyn_bytes);
    for (auto& w : b.warps) {
        w.arrived = 0;
        w.live = 0;
    }
    for (int t = 0; t < nthreads; ++t) {
        Fiber& f = b.fibers[t];
        f.tid = t;
        f.state = RUNNABLE;
        b.warps[t >> 5].live |= 1u << (t & 31);
        char* top = b.stacks + (size_t)(t + 1) * STACK_BYTES;
        uintptr_t sp = ((uintptr_t)top - 64) & ~(uintptr_t)15;
        void** s = (void**)sp;
        // layout popped by emu_swap: r15 r14 r13 r12 rbx rbp ret
        s[0] = s[1] = s[2] = s[3] = s[4] = s[5] = nullptr;
        s[6] = (void*)&fiber_main;
        s[7] = nullptr;
        f.sp = (void*)sp;
    }
}

static void finish_block(Block& b) {
    for (auto& v : b.shared) free(v.ptr);
    b.shared.clear();
    free(b.dyn);
    b.dyn = nullptr;
}

void launch(dim3 grid, dim3 block, size_t dyn_bytes, const std::function<void()>& body) {
    if (grid.y != 1 || grid.z != 1 || block.y != 1 || block.z != 1 || (int)block.x > MAX_THREADS || block.x == 0) {
        fprintf(stderr, "emu: only 1-D launches with <= %d threads are supported\n", MAX_THREADS);
        abort();
    }
    if (grid.x == 0) return;
    if (g_pool_slots < g_window) {
        if (g_pool) munmap(g_pool, (size_t)g_pool_slots * MAX_THREADS * STACK_BYTES);
        g_pool_slots = 8;
        g_pool = (char*)mmap(nullptr, (size_t)g_pool_slots * MAX_THREADS * STACK_BYTES, PROT_READ | PROT_WRITE,
                             MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_pool == MAP_FAILED) {
            perror("emu mmap");
            abort();
        }
    }
    // nested launches are not supported; save globals for safety
    g_body = &body;
    g_dyn_bytes = dyn_bytes;
    blockDim = block;
    gridDim = grid;
    std::vector<Block> resident(g_window);
    std::vector<int> used(g_window, 0);
    unsigned next_bid = 0;
    int n_resident = 0;
    while (true) {
        for (int s = 0; s < g_window && next_bid < grid.x; ++s)
            if (!used[s]) {
                start_block(resident[s], (int)next_bid++, (int)block.x, s);
                used[s] = 1;
                n_resident++;
            }
        if (n_resident == 0) break;
        bool progressed = false;
        int s0 = (int)(rnd() % (uint64_t)g_window);
        for (int k = 0; k < g_window; ++k) {
            int s = (s0 + k) % g_window;
            if (!used[s]) continue;
            Block& b = resident[s];
            int nt = b.nthreads;
            // run one randomly placed warp-sized chunk of fibers per visit, so blocks interleave finely
            int nchunks = (nt + 31) / 32;
            int c0 = (int)(rnd() % (uint64_t)nchunks);
            int span = 1 + (int)(rnd() % (uint64_t)nchunks);
            for (int c = 0; c < span; ++c) {
                int w = (c0 + c) % nchunks;
                for (int t = w * 32; t < nt && t < w * 32 + 32; ++t) {
                    Fiber& f = b.fibers[t];
                    if (f.state != RUNNABLE) continue;
                    g_block = &b;
                    g_fiber = &f;
                    threadIdx.x = (unsigned)t;
                    threadIdx.y = threadIdx.z = 0;
                    blockIdx.x = (unsigned)b.bid;
                    blockIdx.y = blockIdx.z = 0;
                    emu_swap(&g_sched_sp, f.sp);
                    progressed = true;
                }
                // warp rendezvous release
                if (b.warps[w].arrived && release_warp(b, w)) progressed = true;
            }
            if (b.live > 0 && b.at_block == b.live) {
                b.at_block = 0;
                for (auto& f : b.fibers)
                    if (f.state == AT_BLOCK) f.state = RUNNABLE;
                progressed = true;
            }
            // an exiting lane can complete a warp rendezvous of the remaining lanes
            for (int w = 0; w < (int)b.warps.size(); ++w)
                if (b.warps[w].arrived && release_warp(b, w)) progressed = true;
            if (b.live == 0) {
                finish_block(b);
                used[s] = 0;
                n_resident--;
                progressed = true;
            }
        }
        if (!progressed) {
            // nothing runnable in the visited chunks: make sure something is runnable at all
            bool any = false;
            for (int s = 0; s < g_window && !any; ++s)
                if (used[s])
                    for (auto& f : resident[s].fibers)
                        if (f.state == RUNNABLE) {
                            any = true;
                            break;
                        }
            if (!any) {
                fprintf(stderr, "emu: deadlock (all threads blocked at barriers that cannot complete)\n");
                abort();
            }
        }
    }
}

}  // namespace emu

// ---- guarded device allocations ---------------------------------------------------------------
#include <sys/mman.h>
#include <unistd.h>
#include <map>
#include <mutex>
namespace {
std::mutex g_alloc_mu;
std::map<void*, std::pair<void*, size_t>> g_allocs;  // user pointer -> (mapping base, mapping bytes)
}  // namespace

void* emu_device_alloc(size_t n) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t body = (n + page - 1) / page * page;
    const size_t total = body + 2 * page;
    char* base = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == (char*)MAP_FAILED) return nullptr;
    mprotect(base, page, PROT_NONE);
    mprotect(base + page + body, page, PROT_NONE);
    char* user = base + page + ((body - n) & ~(size_t)255);  // 256-byte aligned like cudaMalloc
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    g_allocs[user] = {base, total};
    return user;
}

void emu_device_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) return;
    munmap(it->second.first, it->second.second);
    g_allocs.erase(it);
}
