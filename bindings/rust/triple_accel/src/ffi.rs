//! Raw bindings of include/triple_accel_b200.h (ABI version 2).
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Copy, Clone)]
pub struct TaCosts {
    pub mismatch: u8,
    pub gap: u8,
    pub start_gap: u8,
    pub transpose: u8, // 0 = None
}
#[repr(C)]
#[derive(Copy, Clone)]
pub struct TaMatch {
    pub start: u64,
    pub end: u64,
    pub k: u32,
    pub _pad: u32,
}
#[repr(C)]
#[derive(Copy, Clone)]
pub struct TaEdit {
    pub edit: u32,
    pub count: u32,
}
pub enum TaCtx {}

pub const TA_NONE: u32 = 0xFFFF_FFFF;
pub const TA_OK: c_int = 0;
pub const TA_ERR_LEN_MISMATCH: c_int = -2;
pub const TA_ERR_BAD_COSTS: c_int = -3;
pub const TA_ERR_NUL_BYTE: c_int = -7;

extern "C" {
    pub fn ta_abi_version() -> c_int;
    pub fn ta_init(device: c_int, out: *mut *mut TaCtx) -> c_int;
    /// one context over several GPUs: every host-buffer batch call is split inside the library
    pub fn ta_init_multi(devices: *const c_int, n_devices: c_int, out: *mut *mut TaCtx) -> c_int;
    pub fn ta_device_count(ctx: *mut TaCtx) -> c_int;
    pub fn ta_multi_uses_nccl(ctx: *mut TaCtx) -> c_int;
    pub fn ta_trim();
    /// 1 = the pairs' lengths vary (tile-ordered kernel), 0 = equal lengths, -1 = default; never changes results
    pub fn ta_set_length_hint(ctx: *mut TaCtx, ragged: c_int) -> c_int;
    pub fn ta_shutdown(ctx: *mut TaCtx);
    pub fn ta_strerror(code: c_int) -> *const c_char;
    pub fn ta_last_error(ctx: *mut TaCtx) -> *const c_char;
    pub fn ta_free(p: *mut c_void);
    pub fn ta_search_default_k(needle_len: usize) -> u32;
    pub fn ta_hamming_batch(ctx: *mut TaCtx, a: *const u8, a_off: *const u64, b: *const u8, b_off: *const u64,
                            n: usize, out: *mut u32) -> c_int;
    pub fn ta_levenshtein_k_batch(ctx: *mut TaCtx, a: *const u8, a_off: *const u64, b: *const u8, b_off: *const u64,
                                  n: usize, k: u32, costs: TaCosts, out: *mut u32) -> c_int;
    pub fn ta_levenshtein_k_trace_batch(ctx: *mut TaCtx, a: *const u8, a_off: *const u64, b: *const u8,
                                        b_off: *const u64, n: usize, k: u32, costs: TaCosts, out_dist: *mut u32,
                                        out_edits: *mut *mut TaEdit, out_edit_off: *mut *mut u64) -> c_int;
    pub fn ta_levenshtein_exp_batch(ctx: *mut TaCtx, a: *const u8, a_off: *const u64, b: *const u8,
                                    b_off: *const u64, n: usize, costs: TaCosts, out: *mut u32) -> c_int;
    pub fn ta_levenshtein_exp_trace_batch(ctx: *mut TaCtx, a: *const u8, a_off: *const u64, b: *const u8,
                                          b_off: *const u64, n: usize, costs: TaCosts, out_dist: *mut u32,
                                          out_edits: *mut *mut TaEdit, out_edit_off: *mut *mut u64) -> c_int;
    pub fn ta_levenshtein_search_batch(ctx: *mut TaCtx, needle: *const u8, needle_len: usize, hay: *const u8,
                                       hay_off: *const u64, n: usize, k: u32, search_type: c_int, costs: TaCosts,
                                       anchored: c_int, out_matches: *mut *mut TaMatch,
                                       out_match_off: *mut *mut u64) -> c_int;
    pub fn ta_hamming_search_naive_batch(ctx: *mut TaCtx, needle: *const u8, needle_len: usize, hay: *const u8,
                                         hay_off: *const u64, n: usize, k: u32, search_type: c_int,
                                         out_matches: *mut *mut TaMatch, out_match_off: *mut *mut u64) -> c_int;
    pub fn ta_hamming_search_batch(ctx: *mut TaCtx, needle: *const u8, needle_len: usize, hay: *const u8,
                                   hay_off: *const u64, n: usize, k: u32, search_type: c_int,
                                   out_matches: *mut *mut TaMatch, out_match_off: *mut *mut u64) -> c_int;
}
