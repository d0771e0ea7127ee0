// sab_pack.cuh -- the `pack` serialisation of a suffix array on the GPU
// (/root/reference/src/packed_sa.rs:17-88: PackedSuffixArray::from_sa / into_sa).
//
// Format: bincode 1.2 little-endian header (magic u32 = "SA4x", length u32, data.len() u64) followed by
// BitPacker4x blocks of 128 values at bits = 32 - clz(length - 1): 4 interleaved lanes, lane l packs
// values in[4k+l] (k = 0..31) LSB-first; output word 4j+l is word j of lane l.  The last block is
// zero-padded and its trailing zero bytes are dropped.  The bitpacking / bincode crates are not part
// of the reference tree: byte parity with them is unpinned (the reference only tests the round trip).
// One thread per 32-bit output word (pack) / per value (unpack): pure streaming, HBM-bound.
#pragma once
#include "sab_context.cuh"

#define SAB_PACK_MAGIC 2016690515u

static inline unsigned sab_pack_bits(u64 length) {  // src/packed_sa.rs:127-129
    u64 x = length ? length - 1 : 0;
    unsigned b = 0;
    while (x) {
        ++b;
        x >>= 1;
    }
    return b;
}

// words: total output words = blocks * 4 * bits.  Values beyond len read as 0 (padding of the last block).
__global__ void __launch_bounds__(256)
pack_blocks_kernel(const u32* __restrict__ sa, u64 len, u32 bits, u64 words, u32* __restrict__ out) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= words) return;
    const u64 per_block = 4ull * bits;
    const u64 blk = t / per_block;
    const u32 r = (u32)(t - blk * per_block);
    const u32 j = r >> 2, lane = r & 3u;  // word j of lane `lane`
    const u64 base = blk * 128 + lane;
    const u32 lo_bit = j * 32u;
    u32 k = lo_bit / bits;  // first value overlapping this word
    u32 w = 0;
    for (; k < 32u && k * bits < lo_bit + 32u; ++k) {
        const u64 idx = base + 4ull * k;
        u64 v = idx < len ? sa[idx] : 0u;
        if (bits < 32u) v &= (1ull << bits) - 1ull;
        const int sh = (int)(k * bits) - (int)lo_bit;
        w |= sh >= 0 ? (u32)(v << sh) : (u32)(v >> (-sh));
    }
    out[t] = w;
}

// data: packed words (zero-extended past the trimmed tail by the caller)
__global__ void __launch_bounds__(256)
unpack_blocks_kernel(const u32* __restrict__ data, u64 len, u32 bits, u32* __restrict__ sa) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    if (bits == 0) {
        sa[i] = 0;
        return;
    }
    const u64 blk = i >> 7;
    const u32 r = (u32)(i & 127u), k = r >> 2, lane = r & 3u;
    const u32 pos = k * bits, word = pos >> 5, sh = pos & 31u;
    const u32* w = data + blk * 4ull * bits;
    u64 v = w[4u * word + lane] >> sh;
    if (sh + bits > 32u) v |= (u64)w[4u * (word + 1) + lane] << (32u - sh);
    if (bits < 32u) v &= (1ull << bits) - 1ull;
    sa[i] = (u32)v;
}
