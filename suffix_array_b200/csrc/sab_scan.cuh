// sab_scan.cuh -- single-pass chained scan with decoupled look-back (Merrill & Garland) for small
// POD states.  Each tile publishes its aggregate (PARTIAL) as soon as it is known and its
// inclusive prefix once the look-back has resolved; a warp inspects 32 predecessors per step.
//
// The slot tags carry a launch epoch so the status array never needs clearing between launches:
// a slot whose epoch differs from the current launch's reads as EMPTY.
#pragma once
#include "sab_common.cuh"

enum : u32 { SCAN_EMPTY = 0u, SCAN_PARTIAL = 1u, SCAN_INCLUSIVE = 2u };

// One status slot per tile: the state tag and the value travel in ONE aligned 16-byte access, so a
// look-back step costs a single L2 round trip and needs no fence between "flag" and "value"
// (16-byte accesses of one thread to an aligned address are performed as one transaction).
struct alignas(16) ScanSlot {
    u32 tag;   // (epoch << 2) | state
    u32 w[3];  // the POD value (at most 12 bytes)
};

template <typename T>
struct TileState {
    ScanSlot* slots;  // [tiles]
    u32 epoch;        // 1 .. 2^30-1, unique per launch
};

__device__ __forceinline__ ScanSlot ld_slot(const ScanSlot* p) {
    ScanSlot s;
#ifdef SAB_EMU
    memcpy(&s, p, sizeof(s));  // fibers switch only at explicit yield points
#else
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(s.tag), "=r"(s.w[0]), "=r"(s.w[1]), "=r"(s.w[2])
                 : "l"(p)
                 : "memory");
#endif
    return s;
}
__device__ __forceinline__ void st_slot(ScanSlot* p, const ScanSlot& s) {
#ifdef SAB_EMU
    memcpy(p, &s, sizeof(s));
#else
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(s.tag), "r"(s.w[0]), "r"(s.w[1]), "r"(s.w[2])
                 : "memory");
#endif
}
template <typename T>
__device__ __forceinline__ ScanSlot make_slot(u32 tag, const T& v) {
    static_assert(sizeof(T) % 4 == 0 && sizeof(T) <= 12, "POD scan state: 4, 8 or 12 bytes");
    union {
        T t;
        u32 w[3];
    } u;
    u.w[0] = u.w[1] = u.w[2] = 0;
    u.t = v;
    ScanSlot s;
    s.tag = tag;
    s.w[0] = u.w[0];
    s.w[1] = u.w[1];
    s.w[2] = u.w[2];
    return s;
}
template <typename T>
__device__ __forceinline__ T slot_value(const ScanSlot& s) {
    union {
        T t;
        u32 w[3];
    } u;
    u.w[0] = s.w[0];
    u.w[1] = s.w[1];
    u.w[2] = s.w[2];
    return u.t;
}

template <typename T>
__device__ __forceinline__ T shfl_xor_pod(T v, int m) {
    static_assert(sizeof(T) % 4 == 0, "POD scan state must be a multiple of 4 bytes");
    constexpr int W = sizeof(T) / 4;
    union {
        T t;
        u32 w[W];
    } u;
    u.t = v;
#pragma unroll
    for (int i = 0; i < W; ++i) u.w[i] = __shfl_xor_sync(SAB_FULL, u.w[i], m);
    return u.t;
}

template <typename T, typename Op>
__device__ __forceinline__ T warp_reduce_pod(T v, Op op) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = op(v, shfl_xor_pod(v, m));
    return v;
}

// Look-back of one tile by ONE full warp: publishes `aggregate` (valid in lane 0) as the tile's PARTIAL, combines the
// aggregates of all tiles < `tile` (op commutative + associative), publishes the INCLUSIVE value and returns the
// exclusive prefix in every lane of the warp.  No block barrier inside: the other warps of the block may work on.
// Tiles must be numbered in launch order (a tile may only wait on tiles whose blocks have already started):
// blockIdx.x of a 1-D grid.
template <typename T, typename Op>
__device__ __forceinline__ T tile_exclusive_prefix_warp(const TileState<T>& st, u32 tile, T aggregate, Op op, T identity) {
    const u32 lane = lane_id();
    union {
        T t;
        u32 w[sizeof(T) / 4];
    } bc;
    bc.t = aggregate;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); ++i) bc.w[i] = __shfl_sync(SAB_FULL, bc.w[i], 0);
    aggregate = bc.t;
    const u32 tag = st.epoch << 2;
    if (tile == 0) {
        if (lane == 0) st_slot(&st.slots[0], make_slot<T>(tag | SCAN_INCLUSIVE, aggregate));
        return identity;
    }
    if (lane == 0) st_slot(&st.slots[tile], make_slot<T>(tag | SCAN_PARTIAL, aggregate));
    T running = identity;
    i64 base = (i64)tile - 1;
    while (true) {
        const i64 t = base - (i64)lane;
        u32 f = SCAN_INCLUSIVE;  // virtual tiles before tile 0: inclusive identity
        T val = identity;
        if (t >= 0) {
            while (true) {
                const ScanSlot s = ld_slot(&st.slots[t]);
                f = ((s.tag >> 2) == st.epoch) ? (s.tag & 3u) : (u32)SCAN_EMPTY;
                if (f != SCAN_EMPTY) {
                    val = slot_value<T>(s);
                    break;
                }
                SAB_SPIN_PAUSE();
            }
        }
        const u32 incl = __ballot_sync(SAB_FULL, f == SCAN_INCLUSIVE);
        const u32 first = incl ? (u32)(__ffs((int)incl) - 1) : 31u;
        if (lane > first) val = identity;
        running = op(running, warp_reduce_pod(val, op));
        if (incl) break;
        base -= 32;
    }
    if (lane == 0) st_slot(&st.slots[tile], make_slot<T>(tag | SCAN_INCLUSIVE, op(running, aggregate)));
    return running;
}

// Returns, in every thread of the block, the combination (op, commutative + associative) of the
// aggregates of all tiles < `tile`.  `aggregate` must be valid in thread 0.  Must be called by all
// threads of the block (it contains __syncthreads).  Tile numbering as above.
template <typename T, typename Op>
__device__ __forceinline__ T tile_exclusive_prefix(const TileState<T>& st, u32 tile, T aggregate, Op op, T identity) {
    SAB_SHARED_VAR(T, s_prefix);
    if (warp_id() == 0) {
        const T r = tile_exclusive_prefix_warp<T, Op>(st, tile, aggregate, op, identity);
        if (lane_id() == 0) s_prefix = r;
    }
    __syncthreads();
    T r = s_prefix;
    __syncthreads();  // s_prefix may be reused by a later call in the same kernel
    return r;
}
