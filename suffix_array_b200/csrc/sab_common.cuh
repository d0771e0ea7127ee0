// sab_common.cuh -- shared definitions of the sab200 CUDA engine (sm_100a only).
//
// The same sources also compile under the SIMT emulator of tests/emu (SAB_EMU), which exists only
// so kernel logic can be unit-tested on a machine without a GPU; the product build is nvcc.
#pragma once
#ifndef SAB_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

#define SAB_OK 0
#define SAB_ERR_ARGS (-1)
#define SAB_ERR_OOM (-2)
#define SAB_ERR_CUDA (-3)
#define SAB_ERR_NCCL (-4)
#define SAB_ERR_INTERNAL (-5)

// ---------------------------------------------------------------- launch / shared-memory macros
#ifdef SAB_EMU
#define SAB_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                     \
        (void)(stream);                                                                      \
        emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); }); \
    } while (0)
#define SAB_SHARED_ARRAY(T, name, N) SAB_EMU_SHARED_ARRAY(T, name, N)
#define SAB_SHARED_VAR(T, name) SAB_EMU_SHARED_VAR(T, name)
#define SAB_DYN_SMEM(name) unsigned char* name = (unsigned char*)emu::dyn_smem()
#define SAB_SPIN_PAUSE() emu::yield_spin()
#define SAB_KERNEL_NAME(...) __VA_ARGS__
#else
#define SAB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<dim3(grid), dim3(block), (size_t)(smem), (stream)>>>(__VA_ARGS__)
#define SAB_SHARED_ARRAY(T, name, N) __shared__ T name[N]
#define SAB_SHARED_VAR(T, name) __shared__ T name
#define SAB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define SAB_SPIN_PAUSE() ((void)0)
#define SAB_KERNEL_NAME(...) __VA_ARGS__
#endif

#define SAB_FULL 0xffffffffu

// ---------------------------------------------------------------- memory-order helpers (PTX)
__device__ __forceinline__ u32 ld_acquire_u32(const u32* p) {
#ifdef SAB_EMU
    return *(const volatile u32*)p;
#else
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_release_u32(u32* p, u32 v) {
#ifdef SAB_EMU
    *(volatile u32*)p = v;
#else
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ u32 ld_relaxed_u32(const u32* p) {
#ifdef SAB_EMU
    return *(const volatile u32*)p;
#else
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64* p) {
#ifdef SAB_EMU
    return *(const volatile u64*)p;
#else
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void st_relaxed_u64(u64* p, u64 v) {
#ifdef SAB_EMU
    *(volatile u64*)p = v;
#else
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}

// all earlier memory accesses of the thread (loads included) are performed before its later ones, at GPU scope
__device__ __forceinline__ void fence_acq_rel_gpu() {
#ifndef SAB_EMU
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ u32 lanemask_lt() { return (1u << (threadIdx.x & 31u)) - 1u; }

// warp-wide inclusive scans / reductions over all 32 lanes
__device__ __forceinline__ u32 warp_incl_sum(u32 v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 o = __shfl_up_sync(SAB_FULL, v, d);
        if ((int)lane_id() >= d) v += o;
    }
    return v;
}
__device__ __forceinline__ u32 warp_incl_max(u32 v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 o = __shfl_up_sync(SAB_FULL, v, d);
        if ((int)lane_id() >= d) v = v > o ? v : o;
    }
    return v;
}

static inline u64 div_up64(u64 a, u64 b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- host-side error plumbing
void sab_set_error(const char* fmt, ...);
#define SAB_CUDA_TRY(expr)                                                                  \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            sab_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? SAB_ERR_OOM : SAB_ERR_CUDA;          \
        }                                                                                   \
    } while (0)
#define SAB_TRY(expr)            \
    do {                         \
        int _rc = (expr);        \
        if (_rc != SAB_OK) return _rc; \
    } while (0)
#define SAB_LAUNCH_CHECK() SAB_CUDA_TRY(cudaGetLastError())
