//! Raw declarations of `include/svfm.h`.  NOT COMPILED IN THE BUILD IMAGE (no Rust toolchain there).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const SVFM_OK: c_int = 0;
pub const SVFM_ERR_INVALID_FORMAT: c_int = 1;
pub const SVFM_ERR_BLOB_SIZE: c_int = 2;
pub const SVFM_ERR_EMPTY_PATTERN: c_int = 21;
pub const SVFM_ERR_BAD_SYMBOL: c_int = 24;
pub const SVFM_ERR_CAPACITY: c_int = 25;
pub const SVFM_ERR_CUDA: c_int = 30;
pub const SVFM_REVERSED: u32 = 1;
pub const SVFM_SORTED: u32 = 2;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct svfm_type {
    pub pos_bits: u32, // Position: 32 | 64
    pub planes: u32,   // Block2..Block6
    pub vec_bits: u32, // Vector: 32 | 64 | 128
    pub encoder: u32,  // 0 PassThrough | 1 EncodingTable
}
#[repr(C)]
pub struct svfm_index { _private: [u8; 0] }

extern "C" {
    pub fn svfm_load(blob: *const u8, blob_len: usize, t: svfm_type, device: c_int,
                     out: *mut *mut svfm_index, err_detail: *mut u64) -> c_int;
    pub fn svfm_free(ix: *mut svfm_index);
    pub fn svfm_count_batch(ix: *mut svfm_index, pats: *const u8, offs: *const u64, n: u64, fixed_len: u32,
                            flags: u32, counts_out: *mut c_void) -> c_int;
    pub fn svfm_locate_batch_alloc(ix: *mut svfm_index, pats: *const u8, offs: *const u64, n: u64, fixed_len: u32,
                                   flags: u32, out_offs: *mut u64, positions: *mut *mut c_void,
                                   total: *mut u64) -> c_int;
    pub fn svfm_free_positions(positions: *mut c_void);
    pub fn svfm_last_error() -> *const c_char;
}
