/*
 * sab200.h -- C ABI of the B200-native suffix-array engine (libsab200.so).
 *
 * This is the drop-in boundary for the hot path of the Rust crate hucsmn/suffix_array v0.5.0:
 * every entry point names the reference interface it replaces (file:line under /root/reference).
 * Plain pointers and sizes only; the caller owns every host buffer; the library owns only device
 * memory, streams and events, and keeps no host pointer after a call returns.  There is no CPU
 * fallback: without a CUDA device every compute entry point returns SAB200_ERR_CUDA.
 *
 * Return codes: 0 ok; -1 bad arguments; -2 host/device out of memory; -3 CUDA error;
 * -4 NCCL error; -5 internal error.  sab200_last_error() describes the last failure of the
 * calling thread's most recent call (process-wide string, best effort).
 */
#ifndef SAB200_H
#define SAB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAB200_OK 0
#define SAB200_ERR_ARGS (-1)
#define SAB200_ERR_OOM (-2)
#define SAB200_ERR_CUDA (-3)
#define SAB200_ERR_NCCL (-4)
#define SAB200_ERR_INTERNAL (-5)

/* Replaces `pub const MAX_LENGTH` (src/saca.rs:6, re-exported src/lib.rs:53).  The reference's
 * i32::MAX came from divsufsort's signed indices; the u32 suffix array and the u32 bucket prefix
 * sums (src/sa.rs:112-116) allow n + 1 <= u32::MAX. */
#define SAB200_MAX_LENGTH 0xFFFFFFFEull

/* Length of the bucket table of enable_buckets (src/sa.rs:95): 256 * 257 + 1. */
#define SAB200_BKT_LEN 65793u

#define SAB200_MAX_ROUNDS 64

/* Counters of the last construction on the calling process (for the roofline report). */
typedef struct sab200_stats {
    uint64_t n;
    uint32_t sigma;
    uint32_t bits_per_symbol;
    uint32_t symbols_per_key;
    uint32_t rounds;
    uint64_t active[SAB200_MAX_ROUNDS];
    uint32_t passes[SAB200_MAX_ROUNDS];
    uint64_t radix_pass_launches;
    uint64_t radix_pass_records;
    uint64_t radix_pass_bytes;   /* sum of 2*(K+V)*m over the radix-pass launches */
    double radix_pass_ms;        /* sum of their CUDA-event durations (profiling on) */
    double hist_ms;
    double pack_ms;
    double rank_ms;
    double gather_ms;
    double total_ms;
    double h2d_ms, d2h_ms;
    uint64_t kernel_launches;
} sab200_stats;

/* ---- construction -------------------------------------------------------------------------
 * Replaces saca::saca (src/saca.rs:9-15), the only callee of SuffixArray::new / set
 * (src/sa.rs:25,32): fills all n+1 entries of `sa`, sa[0] = n (src/saca.rs:13), sa[1..] = the
 * suffix starts in increasing suffix order (what cdivsufsort::sort_in_place wrote, src/saca.rs:14).
 * `s` (n bytes) and `sa` (n+1 entries) are HOST buffers owned by the caller (src/sa.rs:24).
 * ngpus: 1 (other values are accepted only when the multi-GPU path is built).
 * The reference panics when n > MAX_LENGTH (src/saca.rs:10); this returns SAB200_ERR_ARGS. */
int32_t sab200_saca(const uint8_t* s, uint64_t n, uint32_t* sa, int32_t ngpus);

/* Same computation with DEVICE buffers already resident on `device` (no copies): d_s holds n
 * bytes, d_sa receives n+1 entries.  Used to time the device pipeline without PCIe. */
int32_t sab200_saca_device(const uint8_t* d_s, uint64_t n, uint32_t* d_sa, int32_t device);

/* ---- introspection ------------------------------------------------------------------------ */
int32_t sab200_get_stats(sab200_stats* out);
void sab200_set_profiling(int32_t on); /* per-launch CUDA events for the stats above */
const char* sab200_last_error(void);
int32_t sab200_device_count(void);     /* number of CUDA devices visible; 0 if none */
const char* sab200_version(void);
void sab200_shutdown(void);            /* releases device memory, streams and events */

#ifdef __cplusplus
}
#endif
#endif /* SAB200_H */
