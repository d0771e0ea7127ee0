// sab_scan.cuh -- single-pass chained scan with decoupled look-back (Merrill & Garland) for small
// POD states.  Each tile publishes its aggregate (PARTIAL) as soon as it is known and its
// inclusive prefix once the look-back has resolved; a warp inspects 32 predecessors per step.
//
// The flag words carry a launch epoch so the status arrays never need clearing between launches:
// a word whose epoch differs from the current launch's reads as EMPTY.
#pragma once
#include "sab_common.cuh"

enum : u32 { SCAN_EMPTY = 0u, SCAN_PARTIAL = 1u, SCAN_INCLUSIVE = 2u };

template <typename T>
struct TileState {
    u32* flags;    // [tiles]   (epoch << 2) | state
    T* partial;    // [tiles]
    T* inclusive;  // [tiles]
    u32 epoch;     // 1 .. 2^30-1, unique per launch
};

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

// Returns, in every thread of the block, the combination (op, commutative + associative) of the
// aggregates of all tiles < `tile`.  `aggregate` must be valid in thread 0.  Must be called by all
// threads of the block (it contains __syncthreads).  Tiles must be numbered in launch order (a
// tile may only wait on tiles whose blocks have already started): blockIdx.x of a 1-D grid.
template <typename T, typename Op>
__device__ __forceinline__ T tile_exclusive_prefix(const TileState<T>& st, u32 tile, T aggregate, Op op, T identity) {
    SAB_SHARED_VAR(T, s_prefix);
    if (warp_id() == 0) {
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
            if (lane == 0) {
                st.inclusive[0] = aggregate;
                st_release_u32(&st.flags[0], tag | SCAN_INCLUSIVE);
                s_prefix = identity;
            }
        } else {
            if (lane == 0) {
                st.partial[tile] = aggregate;
                st_release_u32(&st.flags[tile], tag | SCAN_PARTIAL);
            }
            T running = identity;
            i64 base = (i64)tile - 1;
            while (true) {
                const i64 t = base - (i64)lane;
                u32 f = SCAN_INCLUSIVE;  // virtual tiles before tile 0: inclusive identity
                T val = identity;
                if (t >= 0) {
                    while (true) {
                        const u32 wv = ld_acquire_u32(&st.flags[t]);
                        f = ((wv >> 2) == st.epoch) ? (wv & 3u) : (u32)SCAN_EMPTY;
                        if (f != SCAN_EMPTY) break;
                        SAB_SPIN_PAUSE();
                    }
                    val = (f == SCAN_INCLUSIVE) ? st.inclusive[t] : st.partial[t];
                }
                const u32 incl = __ballot_sync(SAB_FULL, f == SCAN_INCLUSIVE);
                const u32 first = incl ? (u32)(__ffs((int)incl) - 1) : 31u;
                if (lane > first) val = identity;
                running = op(running, warp_reduce_pod(val, op));
                if (incl) break;
                base -= 32;
            }
            if (lane == 0) {
                st.inclusive[tile] = op(running, aggregate);
                st_release_u32(&st.flags[tile], tag | SCAN_INCLUSIVE);
                s_prefix = running;
            }
        }
    }
    __syncthreads();
    T r = s_prefix;
    __syncthreads();  // s_prefix may be reused by a later call in the same kernel
    return r;
}
