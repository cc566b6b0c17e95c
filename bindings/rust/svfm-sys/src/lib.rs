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
pub const SVFM_OFFS32: u32 = 4;
// svfm_set_tuning keys (results never depend on them)
pub const SVFM_TUNE_SORT_MIN: c_int = 0;
pub const SVFM_TUNE_CHUNK: c_int = 1;
pub const SVFM_TUNE_SWEEP_MIN: c_int = 2;
pub const SVFM_TUNE_EXT_BITS: c_int = 3;
pub const SVFM_TUNE_WORKERS: c_int = 4;
pub const SVFM_TUNE_ILV: c_int = 5;
pub const SVFM_TUNE_BUCKET_SORTBACK: c_int = 6;
pub const SVFM_TUNE_SMALL_MAX: c_int = 7;
pub const SVFM_TUNE_TEXT: c_int = 8;
pub const SVFM_TUNE_L2_PERSIST: c_int = 9;
pub const SVFM_TUNE_OWN_RADIX: c_int = 10;
pub const SVFM_TUNE_FULL_SA: c_int = 11;
pub const SVFM_TUNE_SWEEP_OCC: c_int = 12;
pub const SVFM_TUNE_AUTO: u64 = 0xffff_ffff_ffff_fffe;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct svfm_type {
    pub pos_bits: u32, // Position: 32 | 64
    pub planes: u32,   // Block2..Block6
    pub vec_bits: u32, // Vector: 32 | 64 | 128
    pub encoder: u32,  // 0 PassThrough | 1 EncodingTable
}
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct svfm_info {
    pub type_: svfm_type,
    pub device: i32,
    pub symbol_count: u32,
    pub kmer_size: u32,
    pub sampling_ratio: u32,
    pub text_len: u64,
    pub suffix_array_len: u64,
    pub blocks_len: u64,
    pub sentinel_index: u64,
    pub blob_len: u64,
    pub header_size: u64,
    pub off_suffix_array: u64,
    pub off_rank_checkpoints: u64,
    pub off_blocks: u64,
}
#[repr(C)]
pub struct svfm_index { _private: [u8; 0] }
#[repr(C)]
pub struct svfm_session { _private: [u8; 0] }

extern "C" {
    pub fn svfm_load(blob: *const u8, blob_len: usize, t: svfm_type, device: c_int,
                     out: *mut *mut svfm_index, err_detail: *mut u64) -> c_int;
    pub fn svfm_free(ix: *mut svfm_index);
    pub fn svfm_count_batch(ix: *mut svfm_index, pats: *const u8, offs: *const u64, n: u64, fixed_len: u32,
                            flags: u32, counts_out: *mut c_void) -> c_int;
    pub fn svfm_locate_batch_alloc(ix: *mut svfm_index, pats: *const u8, offs: *const u64, n: u64, fixed_len: u32,
                                   flags: u32, out_offs: *mut c_void /* u64[n+1]; u32[n+1] with SVFM_OFFS32 */,
                                   positions: *mut *mut c_void, total: *mut u64) -> c_int;
    pub fn svfm_free_positions(positions: *mut c_void);
    pub fn svfm_last_error() -> *const c_char;
    // the rest of include/svfm.h
    pub fn svfm_load_device(d_blob: *const u8, blob_len: usize, t: svfm_type, device: c_int,
                            out: *mut *mut svfm_index, err_detail: *mut u64) -> c_int;
    pub fn svfm_check_blob(blob: *const u8, blob_len: usize, t: svfm_type, out: *mut svfm_info, err_detail: *mut u64) -> c_int;
    pub fn svfm_index_info(ix: *const svfm_index, out: *mut svfm_info) -> c_int;
    pub fn svfm_index_memory(ix: *mut svfm_index, out: *mut u64) -> c_int; // [8]: blob, ext table, interleaved occ, scratch, text copy, expanded SA, sweep occ copy, reserved
    pub fn svfm_locate_batch(ix: *mut svfm_index, pats: *const u8, offs: *const u64, n: u64, fixed_len: u32, flags: u32,
                             out_offs: *mut c_void, positions: *mut c_void, capacity: u64, total: *mut u64) -> c_int;
    // packed fixed-length batches: ceil(len*bits/8) bytes per pattern, symbol indices, first symbol in the lowest bits
    pub fn svfm_count_batch_packed(ix: *mut svfm_index, packed: *const u8, n: u64, len: u32, bits: u32, flags: u32,
                                   counts_out: *mut c_void) -> c_int;
    pub fn svfm_locate_batch_packed(ix: *mut svfm_index, packed: *const u8, n: u64, len: u32, bits: u32, flags: u32,
                                    out_offs: *mut c_void, positions: *mut c_void, capacity: u64, total: *mut u64) -> c_int;
    pub fn svfm_pack_patterns(pats: *const u8, n: u64, len: u32, table256: *const u8, bits: u32, packed_out: *mut u8) -> c_int;
    pub fn svfm_count(ix: *mut svfm_index, pattern: *const u8, len: u64, flags: u32, count: *mut u64) -> c_int;
    pub fn svfm_locate(ix: *mut svfm_index, pattern: *const u8, len: u64, flags: u32, positions: *mut c_void,
                       capacity: u64, total: *mut u64) -> c_int;
    pub fn svfm_session_create(ix: *mut svfm_index, out: *mut *mut svfm_session) -> c_int;
    pub fn svfm_session_destroy(s: *mut svfm_session);
    pub fn svfm_session_sync(s: *mut svfm_session) -> c_int;
    pub fn svfm_session_stream(s: *mut svfm_session) -> *mut c_void; // cudaStream_t
    pub fn svfm_count_batch_device(s: *mut svfm_session, d_pats: *const u8, d_offs: *const u64, n: u64, fixed_len: u32,
                                   flags: u32, d_counts_out: *mut c_void) -> c_int;
    pub fn svfm_locate_batch_device(s: *mut svfm_session, d_pats: *const u8, d_offs: *const u64, n: u64, fixed_len: u32,
                                    flags: u32, d_out_offs: *mut c_void /* u64[n+1]; u32[n+1] with SVFM_OFFS32 */,
                                    d_positions: *mut *mut c_void, total: *mut u64) -> c_int;
    pub fn svfm_session_set_timing(s: *mut svfm_session, enabled: c_int) -> c_int;
    pub fn svfm_session_get_timing(s: *mut svfm_session, ms: *mut f64, launches: *mut u64, reset: c_int) -> c_int; // [SVFM_PHASE_MAX = 8]
    pub fn svfm_host_alloc(bytes: usize) -> *mut c_void;
    pub fn svfm_host_free(p: *mut c_void);
    pub fn svfm_set_tuning(key: c_int, value: u64) -> c_int;
    pub fn svfm_launch_count() -> u64;
    pub fn svfm_version() -> *const c_char;
    pub fn svfm_blob_size(t: svfm_type, text_len: u64, symbol_count: u32, kmer_size: u32, sampling_ratio: u32,
                          blob_size: *mut u64, err_detail: *mut u64) -> c_int;
    pub fn svfm_build(t: svfm_type, text: *const u8, text_len: u64, symbol_count: u32, table256: *const u8,
                      kmer_size: u32, sampling_ratio: u32, device: c_int, blob_out: *mut u8, blob_len: u64,
                      err_detail: *mut u64) -> c_int;
    pub fn svfm_build_device(t: svfm_type, d_text: *const u8, text_len: u64, symbol_count: u32, table256: *const u8,
                             kmer_size: u32, sampling_ratio: u32, device: c_int, d_blob_out: *mut u8, blob_len: u64,
                             err_detail: *mut u64) -> c_int;
}
