// cuda_emu.h -- a small SIMT emulator so the CUDA kernels of suffix_array_b200/csrc can be
// unit-tested on a machine without a GPU.
//
// TEST INFRASTRUCTURE ONLY.  The product library (libsab200.so) is built by nvcc for sm_100a and
// never includes this header; the package refuses to run without a GPU.  This header is force-
// included (g++ -x c++ -DSAB_EMU -include tests/emu/cuda_emu.h) when building libsab200_emu.so,
// which only tests/test_emu_*.py load.  It runs every CUDA thread as a fiber: a window of
// resident blocks is interleaved in a seeded random order, __syncthreads / warp collectives are
// rendezvous points, spin loops yield.  That exercises the same kernel source -- indexing, scans,
// decoupled look-back (partial AND inclusive paths), stability -- but says nothing about
// performance or about the real memory model.  Parity claims rest on the `-m gpu` tests alone.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(x) __attribute__((aligned(x)))
#define __noinline__ __attribute__((noinline))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct emu_uint3 {
    unsigned x, y, z;
};
struct uint2 {
    unsigned x, y;
};
struct uint4 {
    unsigned x, y, z, w;
};
struct ulonglong2 {
    unsigned long long x, y;
};
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

extern emu_uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
static const int warpSize = 32;

namespace emu {
void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()>& body);
void yield_spin();                 // a thread spinning on global memory lets others run
void sync_block();                 // __syncthreads
void sync_warp(unsigned mask);     // rendezvous of the live lanes named by mask
void* shared_get(const void* key, size_t bytes);
void* dyn_smem();
uint64_t* warp_slots();            // 32 exchange slots of the calling thread's warp
uint32_t warp_live_mask();         // lanes of this warp that have not exited
void set_seed(uint64_t seed);      // scheduler interleaving seed
void set_window(int resident_blocks);
}  // namespace emu

#define SAB_EMU_SHARED_ARRAY(T, name, N) \
    static char name##_emu_key;          \
    T* name = (T*)emu::shared_get(&name##_emu_key, sizeof(T) * (size_t)(N))
#define SAB_EMU_SHARED_VAR(T, name) \
    static char name##_emu_key;     \
    T& name = *(T*)emu::shared_get(&name##_emu_key, sizeof(T))

// ------------------------------------------------------------------ intrinsics
static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned m = 0xffffffffu) { emu::sync_warp(m); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline unsigned __activemask() { return emu::warp_live_mask(); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31));
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }
static inline void __nanosleep(unsigned) { emu::yield_spin(); }

template <typename T>
static inline uint64_t emu_to_bits(T v) {
    static_assert(sizeof(T) <= 8, "emu shuffles carry at most 64 bits");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
static inline T emu_from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
// every collective: deposit, rendezvous, compute, rendezvous
template <typename F>
static inline auto emu_collective(unsigned mask, uint64_t mine, F f) -> decltype(f((const uint64_t*)0, 0u, 0)) {
    uint64_t* slots = emu::warp_slots();
    int lane = (int)(threadIdx.x & 31);
    slots[lane] = mine;
    emu::sync_warp(mask);
    auto r = f((const uint64_t*)slots, emu::warp_live_mask() & mask, lane);
    emu::sync_warp(mask);
    return r;
}
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    return emu_collective(mask, emu_to_bits(v), [&](const uint64_t* s, unsigned, int lane) {
        int base = lane & ~(width - 1);
        return emu_from_bits<T>(s[base + (src & (width - 1))]);
    });
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    return emu_collective(mask, emu_to_bits(v), [&](const uint64_t* s, unsigned, int lane) {
        int base = lane & ~(width - 1);
        int src = lane - (int)d;
        return emu_from_bits<T>(s[src < base ? lane : src]);
    });
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    return emu_collective(mask, emu_to_bits(v), [&](const uint64_t* s, unsigned, int lane) {
        int base = lane & ~(width - 1);
        int src = lane + (int)d;
        return emu_from_bits<T>(s[src >= base + width ? lane : src]);
    });
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int m, int width = 32) {
    (void)width;
    return emu_collective(mask, emu_to_bits(v), [&](const uint64_t* s, unsigned, int lane) {
        return emu_from_bits<T>(s[(lane ^ m) & 31]);
    });
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    return emu_collective(mask, pred ? 1u : 0u, [&](const uint64_t* s, unsigned live, int) {
        unsigned r = 0;
        for (int i = 0; i < 32; ++i)
            if (((live >> i) & 1u) && s[i]) r |= 1u << i;
        return r;
    });
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == (emu::warp_live_mask() & m); }
template <typename T>
static inline unsigned __match_any_sync(unsigned mask, T v) {
    return emu_collective(mask, emu_to_bits(v), [&](const uint64_t* s, unsigned live, int lane) {
        unsigned r = 0;
        for (int i = 0; i < 32; ++i)
            if (((live >> i) & 1u) && s[i] == s[lane]) r |= 1u << i;
        return r;
    });
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    return emu_collective(mask, (uint64_t)v, [&](const uint64_t* s, unsigned live, int) {
        unsigned r = 0;
        for (int i = 0; i < 32; ++i)
            if ((live >> i) & 1u) r += (unsigned)s[i];
        return r;
    });
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    return emu_collective(mask, (uint64_t)v, [&](const uint64_t* s, unsigned live, int) {
        unsigned r = 0;
        for (int i = 0; i < 32; ++i)
            if (((live >> i) & 1u) && (unsigned)s[i] > r) r = (unsigned)s[i];
        return r;
    });
}

// atomics: one OS thread, fibers switch only at explicit points -> plain read-modify-write
template <typename T>
static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, int v) { unsigned o = *p; *p = o + (unsigned)v; return o; }
template <typename T>
static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T>
static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T>
static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T>
static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <typename T>
static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T>
static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// ------------------------------------------------------------------ runtime API subset
typedef int cudaError_t;
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event { double t; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaHostRegisterDefault = 0, cudaEventDefault = 0, cudaEventDisableTiming = 2 };
struct cudaDeviceProp {
    char name[256];
    int multiProcessorCount;
    size_t totalGlobalMem;
    int major, minor;
};
static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : (e == 2 ? "out of memory (emu)" : "error (emu)"); }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
// device allocations end (within 256 B) at an inaccessible guard page: an out-of-bounds kernel access faults
// here like "illegal memory access" on the GPU instead of silently landing in a neighbouring malloc block
void* emu_device_alloc(size_t n);
void emu_device_free(void* p);
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = emu_device_alloc(n ? n : 1);
    if (*p) memset(*p, 0xCD, n);  // poison: device memory is not zero-initialised
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <typename T>
static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { emu_device_free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T>
static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p));
    strcpy(p->name, "SIMT emulator (no GPU)");
    p->multiProcessorCount = 4;
    p->totalGlobalMem = (size_t)8 << 30;
    p->major = 10;
    return cudaSuccess;
}
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = (size_t)8 << 30; *t = (size_t)8 << 30; return cudaSuccess; }
double emu_now_ms();
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event{0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = emu_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
