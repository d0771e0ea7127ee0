//! Drop-in replacement for `src/saca.rs` of hucsmn/suffix_array v0.5.0 (reference lines 1-23).
//!
//! NOT COMPILED IN THIS REPOSITORY: the build image has no Rust toolchain.  This file and
//! `build.rs` / `ffi.rs` next to it are the binding a maintainer of the crate would add; the C ABI
//! they call (include/sab200.h) is exercised by the C++ and Python hosts of this repository.
//!
//! Behaviour kept from the reference: both asserts (src/saca.rs:10-11), the sentinel entry
//! `sa[0] = n` (written by the library), panic on failure (the reference's cdivsufsort wrapper
//! panics when divsufsort returns non-zero).  Changed on purpose: MAX_LENGTH (see below).

use super::ffi;

/// Maximum length of the input string.
///
/// The reference's `i32::MAX` came from divsufsort's signed 32-bit indices (src/saca.rs:6).  The
/// suffix array itself is `Vec<u32>` of n+1 entries and the bucket prefix sums are u32
/// (src/sa.rs:112-116), so the engine accepts n + 1 <= u32::MAX.
pub const MAX_LENGTH: usize = (std::u32::MAX - 1) as usize;

/// Texts below this size stay on one GPU: the exchange steps of the sharded construction do not pay.
const MULTI_GPU_MIN_LEN: usize = 256 << 20;

/// GPUs to use: `SAB200_GPUS` if set, else every visible device for large texts, one otherwise.
fn gpus_for(len: usize) -> i32 {
    if let Some(v) = std::env::var("SAB200_GPUS").ok().and_then(|v| v.parse::<i32>().ok()) {
        return v.max(1);
    }
    let have = unsafe { ffi::sab200_device_count() };
    if len >= MULTI_GPU_MIN_LEN && have > 1 { have.min(16) } else { 1 }
}

/// Wrapper of the underlying suffix array construction algorithm (GPU prefix doubling; with more than
/// one GPU the text is sharded and the library runs its distributed sample sort + NCCL rounds behind this
/// same call -- the signature of the reference's saca() is unchanged).
pub fn saca(s: &[u8], sa: &mut [u32]) {
    assert!(s.len() <= MAX_LENGTH);
    assert_eq!(s.len() + 1, sa.len());

    let rc = unsafe { ffi::sab200_saca(s.as_ptr(), s.len() as u64, sa.as_mut_ptr(), gpus_for(s.len())) };
    assert_eq!(rc, 0, "sab200_saca failed: {}", ffi::last_error());
}
