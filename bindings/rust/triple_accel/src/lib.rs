//! `triple_accel` (v0.4.0 API) served by the B200 batched edit-distance engine.
//!
//! Every public name, argument and return type of the reference crate is kept (reference `src/lib.rs:122-174,
//! 196-235`); the work is done by `libtriple_accel_b200.so` through its C ABI.  Single-pair calls are batches of one;
//! the `*_batch` functions at the end expose the path the GPU is built for.  Errors keep the crate's behaviour:
//! contract violations panic, "not within k" is `None`, and there is no CPU fallback (no usable CUDA device = panic).
//!
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image); kept next to the library as the binding a
//! maintainer would start from.  `tests/test_api_names.py` checks the same name list against the Python mirror.
pub mod ffi;
pub mod hamming;
pub mod levenshtein;

pub use crate::hamming::*;
pub use crate::levenshtein::*;

/// reference `src/lib.rs:134-142`: start inclusive, end exclusive, `k` = cost of the match
#[derive(Debug, PartialEq)]
pub struct Match {
    pub start: usize,
    pub end: usize,
    pub k: u32,
}

/// reference `src/lib.rs:147-154` (the C ABI uses the same order: 0 = Match ... 4 = Transpose)
#[derive(Debug, PartialEq, Copy, Clone)]
pub enum EditType {
    Match,
    Mismatch,
    AGap,
    BGap,
    Transpose,
}

/// reference `src/lib.rs:159-165`
#[derive(Debug, PartialEq)]
pub struct Edit {
    pub edit: EditType,
    pub count: usize,
}

/// reference `src/lib.rs:170-174`
#[derive(Debug, PartialEq, Copy, Clone)]
pub enum SearchType {
    All,
    Best,
}

/// reference `src/lib.rs:196-206`: a zeroed byte string.  (The 16-byte alignment only mattered to the reference's
/// word-wise CPU loops; the GPU path takes any alignment.)
pub fn alloc_str(len: usize) -> Vec<u8> {
    vec![0u8; len]
}

/// reference `src/lib.rs:228-235`
pub fn fill_str(dest: &mut [u8], src: &[u8]) {
    assert!(dest.len() >= src.len());
    dest[..src.len()].copy_from_slice(src);
}

// ---- plumbing shared by the two modules ---------------------------------------------------------------------------

/// One engine context per thread (the crate's functions are re-entrant; so is this).  `TA_DEVICES=0,1,..,7` makes it
/// ONE context over those GPUs (`ta_init_multi`: every batch call is split across them inside the library); else the
/// device is LOCAL_RANK when set (one process per GPU), else 0.
pub(crate) fn ctx() -> *mut ffi::TaCtx {
    thread_local!(static CTX: *mut ffi::TaCtx = unsafe {
        let mut c = std::ptr::null_mut();
        let devs: Vec<i32> = std::env::var("TA_DEVICES").ok()
            .map(|s| s.split(',').filter_map(|x| x.trim().parse().ok()).collect()).unwrap_or_default();
        let dev = std::env::var("LOCAL_RANK").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let rc = if devs.len() > 1 { ffi::ta_init_multi(devs.as_ptr(), devs.len() as i32, &mut c) } else { ffi::ta_init(dev, &mut c) };
        if rc != ffi::TA_OK {
            panic!("triple_accel_b200: ta_init failed ({}): there is no CPU fallback", rc)
        }
        c
    });
    CTX.with(|c| *c)
}

pub(crate) fn check(rc: i32) {
    if rc != ffi::TA_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::ta_strerror(rc)).to_string_lossy().into_owned() };
        panic!("triple_accel_b200: {}", msg)
    }
}

pub(crate) fn edit_type(code: u32) -> EditType {
    match code {
        0 => EditType::Match,
        1 => EditType::Mismatch,
        2 => EditType::AGap,
        3 => EditType::BGap,
        _ => EditType::Transpose,
    }
}

/// Takes ownership of a (matches, offsets) pair returned by a one-haystack search call.
pub(crate) unsafe fn take_matches(m: *mut ffi::TaMatch, mo: *mut u64) -> Vec<Match> {
    let n = *mo.add(1) as usize;
    let v = (0..n)
        .map(|i| {
            let x = *m.add(i);
            Match { start: x.start as usize, end: x.end as usize, k: x.k }
        })
        .collect();
    ffi::ta_free(m as *mut _);
    ffi::ta_free(mo as *mut _);
    v
}
