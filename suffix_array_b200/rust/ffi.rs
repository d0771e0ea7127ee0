//! `extern "C"` declarations of include/sab200.h (libsab200.so).  NOT COMPILED HERE (no Rust toolchain).
#![allow(dead_code)]
use std::ffi::CStr;
use std::os::raw::c_char;

#[repr(C)]
pub struct Sab200Index {
    _private: [u8; 0],
}

#[link(name = "sab200")]
extern "C" {
    pub fn sab200_saca(s: *const u8, n: u64, sa: *mut u32, ngpus: i32) -> i32;
    pub fn sab200_enable_buckets(s: *const u8, n: u64, bkt: *mut u32) -> i32;
    pub fn sab200_check(s: *const u8, n: u64, sa: *const u32, sa_len: u64) -> i32;
    pub fn sab200_index_create(s: *const u8, n: u64, sa: *const u32, sa_len: u64, bkt_or_null: *const u32, ngpus: i32) -> *mut Sab200Index;
    pub fn sab200_index_destroy(ix: *mut Sab200Index);
    pub fn sab200_search_all_batch(ix: *mut Sab200Index, pats: *const u8, offs: *const u64, np: u64, lo: *mut u32, hi: *mut u32) -> i32;
    pub fn sab200_contains_batch(ix: *mut Sab200Index, pats: *const u8, offs: *const u64, np: u64, out: *mut u8) -> i32;
    pub fn sab200_search_lcp_batch(ix: *mut Sab200Index, pats: *const u8, offs: *const u64, np: u64, start: *mut u32, end: *mut u32) -> i32;
    // feature "pack": GPU forms of PackedSuffixArray::from_sa + dump_bytes / load_bytes + into_sa (src/packed_sa.rs)
    pub fn sab200_pack_bound(sa_len: u64) -> u64;
    pub fn sab200_pack(sa: *const u32, sa_len: u64, out: *mut u8, out_cap: u64, out_len: *mut u64) -> i32;
    pub fn sab200_unpack(bytes: *const u8, nbytes: u64, sa: *mut u32, sa_cap: u64, sa_len: *mut u64) -> i32;
    pub fn sab200_device_count() -> i32;
    pub fn sab200_last_error() -> *const c_char;
    pub fn sab200_shutdown();
}

pub fn last_error() -> String {
    unsafe { CStr::from_ptr(sab200_last_error()).to_string_lossy().into_owned() }
}
