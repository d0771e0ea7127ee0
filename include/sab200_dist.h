/*
 * sab200_dist.h -- per-rank step functions of the multi-GPU construction (libsab200.so).
 *
 * The reference has nothing distributed; this is the sharded form of the same
 * saca() (/root/reference/src/saca.rs:9-15) for texts spread over the GPUs of one box.  One process
 * per GPU; the exchange steps between these calls are collectives issued by the host driver
 * (suffix_array_b200/dist.py: torch.distributed all_to_all_single / all_gather / all_reduce over
 * NCCL on NVLink).  Every pointer below is a DEVICE pointer on `device` unless marked host; every
 * call synchronises the library stream before returning.  Return codes as in sab200.h.
 *
 * Layout: rank g of P owns text positions [g*B, min((g+1)*B, n)) and their rank[] entries
 * (B = ceil(n/P)); after the key exchange it owns a contiguous slice of the suffix array whose
 * groups of equal keys never straddle two GPUs (splitters cut between distinct keys).
 */
#ifndef SAB200_DIST_H
#define SAB200_DIST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 256-bin byte histogram of d_text[0..len) -> d_hist (256 x u64, overwritten). */
int32_t sab200_dist_hist(const uint8_t* d_text, uint64_t len, uint64_t* d_hist, int32_t device);

/* Host-only: from the GLOBAL byte histogram and text length choose the code table (codes 1..sigma,
 * 0 = absent byte), the key radix b = sigma + 1 and the symbols per key k (same cost model as the
 * single-GPU path).  A key is the mixed-radix number of the first k codes; it needs
 * bit_length(b^k - 1) bits. */
int32_t sab200_dist_plan(const uint64_t* hist256, uint64_t n, uint16_t* lut256, int32_t* b, int32_t* k);

/* Keys of `count` consecutive suffixes starting at global position shard_lo.  d_text holds the text
 * from shard_lo on: at least min(count + 64, n - shard_lo) bytes.  d_idx[j] = shard_lo + j. */
int32_t sab200_dist_pack(const uint8_t* d_text, uint64_t shard_lo, uint64_t count, uint64_t n, const uint16_t* lut256,
                         int32_t b, int32_t k, uint64_t* d_keys, uint32_t* d_idx, int32_t device);

/* Destination rank of a key = number of splitters <= key (nsplit = P-1 host values, ascending).
 * counts (host, P x u64) receives how many of the `count` keys go to each rank; the records are
 * stably partitioned by destination into (d_keys_out, d_idx_out). */
int32_t sab200_dist_partition_keys(const uint64_t* d_keys, const uint32_t* d_idx, uint64_t count,
                                   const uint64_t* splitters, int32_t nsplit, uint64_t* d_keys_out,
                                   uint32_t* d_idx_out, uint64_t* counts, int32_t device);

/* Stable LSD radix sort of (key, payload) pairs on bits [0, key_bits).  Buffers 0 hold the input;
 * returns 0 or 1 = which buffer pair holds the result, or < 0 on error. */
int32_t sab200_dist_sort_pairs(uint64_t* d_k0, uint64_t* d_k1, uint32_t* d_v0, uint32_t* d_v1, uint64_t count,
                               int32_t key_bits, int32_t device);

/* After the local sort of this rank's slice (count records, SA positions sa_off .. sa_off+count):
 * d_sa_local[j] = suffix of record j; d_rank_seq[j] = rank of record j (= sa_off + index of its group
 * head); records of groups > 1 are compacted into (d_act_r1, d_act_idx); *n_active (host) = their count. */
int32_t sab200_dist_init_ranks(const uint64_t* d_keys, const uint32_t* d_idx, uint64_t count, uint32_t sa_off,
                               uint32_t* d_sa_local, uint32_t* d_rank_seq, uint32_t* d_act_r1, uint32_t* d_act_idx,
                               uint64_t* n_active, int32_t device);

/* Layout of the distributed rank[] array, named by (B, P, cyc_shift) in the calls below:
 *   cyc_shift < 0  block layout: GPU g owns positions [g*B, (g+1)*B), the last GPU also the tail (position n);
 *                  local slot of position q = q - g*B
 *   cyc_shift >= 0 block-cyclic layout, B = 2^cyc_shift: owner = (q >> cyc_shift) mod P,
 *                  local slot = ((q >> cyc_shift) / P) << cyc_shift | (q mod B).  Every region of the text is
 *                  spread over all GPUs, so the rank requests of a round do not pile up on the owners of the
 *                  region they point into (measured on the mixed 3.9 GiB text, profiles/r01_multi_gpu.md).
 *
 * Stable partition of (key, val) u32 pairs by the owner of text position key + add; keys equal to 0xFFFFFFFF
 * are dropped (they sort behind the last rank and are not counted).  counts (host, P x u64). */
int32_t sab200_dist_partition_owner(const uint32_t* d_key, const uint32_t* d_val, uint64_t count, uint32_t add,
                                    uint32_t B, int32_t P, int32_t cyc_shift, uint32_t* d_key_out, uint32_t* d_val_out,
                                    uint64_t* counts, int32_t device);

/* Stable partition of (pos, val) u32 pairs by the rank whose suffix-array slice holds SA position pos:
 * slice_start (host, P x u32) = first SA position of every rank's slice, ascending; pos equal to
 * 0xFFFFFFFF is dropped.  Used when the active lists have been rebalanced across the GPUs, so a newly
 * unique suffix may belong to another rank's slice.  counts (host, P x u64). */
int32_t sab200_dist_partition_slices(const uint32_t* d_pos, const uint32_t* d_val, uint64_t count,
                                     const uint32_t* slice_start, int32_t P, uint32_t* d_pos_out, uint32_t* d_val_out,
                                     uint64_t* counts, int32_t device);

/* d_rank_local[slot(d_pos[t])] = d_val[t]: ranks arriving at their owner (slot = d_pos[t] - lo under the
 * block layout); also SA entries arriving at the owner of their slice (block layout, lo = slice offset). */
int32_t sab200_dist_scatter(const uint32_t* d_pos, const uint32_t* d_val, uint64_t count, uint32_t lo, uint32_t B,
                            int32_t P, int32_t cyc_shift, uint32_t* d_rank_local, int32_t device);
/* d_out[t] = d_rank_local[slot(d_pos[t] + add)]   (answering rank[i+h] requests) */
int32_t sab200_dist_gather(const uint32_t* d_pos, uint64_t count, uint32_t add, uint32_t lo, uint32_t B, int32_t P,
                           int32_t cyc_shift, const uint32_t* d_rank_local, uint32_t* d_out, int32_t device);
/* d_key64[t] = (d_r1[t] << 32) | d_r2[t] */
int32_t sab200_dist_make_keys(const uint32_t* d_r1, const uint32_t* d_r2, uint64_t count, uint64_t* d_key64,
                              int32_t device);

/* One re-ranking step on this rank's sorted active records (see rerank_kernel): newly unique suffixes
 * are written to d_sa_local[rank - sa_off]; the rest is compacted into (d_out_r1, d_out_idx), *n_kept
 * (host); (d_upd_idx[j], d_upd_r[j]) lists every changed rank (0xFFFFFFFF in d_upd_idx = unchanged).
 * d_set_pos != NULL (rebalanced active lists): nothing is written to d_sa_local; d_set_pos[j] = SA
 * position of record j if it became unique, else 0xFFFFFFFF -- the caller routes (d_set_pos, d_idx)
 * to the owners of the slices (sab200_dist_partition_slices). */
int32_t sab200_dist_rerank(const uint64_t* d_key64, const uint32_t* d_idx, uint64_t m, uint32_t sa_off,
                           uint32_t* d_sa_local, uint32_t* d_out_r1, uint32_t* d_out_idx, uint32_t* d_upd_idx,
                           uint32_t* d_upd_r, uint32_t* d_set_pos, uint64_t* n_kept, int32_t device);

/* Lazy inverse suffix array, distributed (block layout only; SAB_DIST_LAZY=1 in dist.py): only the ranks of
 * active suffixes are stored at their owners, rank[] starts 0xFFFFFFFF (EMPTY).  At the owner, after
 * sab200_dist_gather: lazy_collect compacts the requests answered EMPTY into (key of suffix d_q[t] + h packed
 * from the text shard, t); the keys travel to the GPU whose slice holds them (sab200_dist_partition_keys +
 * all_to_all), lower_bound there turns each into its rank (slice offset + index in the sorted keys;
 * 0xFFFFFFFF if absent, which would be a bug), and lazy_fill stores the returned ranks into the answers and
 * into rank[] (memoised). */
int32_t sab200_dist_lazy_collect(const uint32_t* d_q, const uint32_t* d_ans, uint64_t count, uint32_t h, uint64_t shard_lo,
                                 const uint8_t* d_text, uint64_t n, const uint16_t* lut256, int32_t b, int32_t k,
                                 uint64_t* d_keys_out, uint32_t* d_slot_out, uint64_t* n_unresolved, int32_t device);
int32_t sab200_dist_lower_bound(const uint64_t* d_sorted_keys, uint64_t R, const uint64_t* d_keys, uint64_t count,
                                uint32_t sa_off, uint32_t* d_rank_out, int32_t device);
int32_t sab200_dist_lazy_fill(const uint32_t* d_slot, const uint32_t* d_rank, uint64_t count, const uint32_t* d_q, uint32_t h,
                              uint32_t lo, uint32_t* d_ans, uint32_t* d_rank_local, int32_t device);

/* Peer-to-peer form of the round exchanges (NVLink): peer_rank_ptrs (host array of P device addresses)
 * are the rank[] arrays of all GPUs mapped into this process (symmetric memory; layout as above).
 *   gather:  d_key64[t] = (d_r1[t] << 32) | rank[d_idx[t] + h], loaded from the owner GPU
 *   scatter: rank[d_idx[t]] = d_val[t] stored into the owner GPU (d_idx[t] == 0xFFFFFFFF: skip)
 * The caller orders the phases across ranks (nobody writes while anyone still reads). */
int32_t sab200_dist_gather_p2p(const uint32_t* d_r1, const uint32_t* d_idx, uint64_t m, uint32_t h, uint32_t B, int32_t P,
                               int32_t cyc_shift, const uint64_t* peer_rank_ptrs, uint64_t* d_key64, int32_t device);
int32_t sab200_dist_scatter_p2p(const uint32_t* d_idx, const uint32_t* d_val, uint64_t count, uint32_t B, int32_t P,
                                int32_t cyc_shift, const uint64_t* peer_rank_ptrs, int32_t device);

/* Key exchange fused into the partition kernel.  count_keys: counts (host, P x u64) per destination.
 * partition_keys_p2p: the partition pass stores destination d's records straight into GPU d's receive
 * buffers (peer_key_ptrs[d] / peer_idx_ptrs[d], mapped device addresses) from record offsets[d] on --
 * no staging copy, no all_to_all; the caller barriers before the buffers are read. */
int32_t sab200_dist_count_keys(const uint64_t* d_keys, uint64_t count, const uint64_t* splitters, int32_t nsplit,
                               uint64_t* counts, int32_t device);
int32_t sab200_dist_partition_keys_p2p(const uint64_t* d_keys, const uint32_t* d_idx, uint64_t count,
                                       const uint64_t* splitters, int32_t nsplit, const uint64_t* offsets,
                                       const uint64_t* peer_key_ptrs, const uint64_t* peer_idx_ptrs, int32_t device);

/* Bracket one construction on this rank: begin() clears the counters of sab200_get_stats() (and arms
 * per-launch event timing when sab200_set_profiling(1)); end() publishes them. */
int32_t sab200_dist_begin(int32_t device);
int32_t sab200_dist_end(int32_t device);

#ifdef __cplusplus
}
#endif
#endif /* SAB200_DIST_H */
