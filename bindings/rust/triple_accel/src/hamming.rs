//! reference `src/hamming.rs`: every Hamming entry point, as names for the one GPU path.
use crate::{check, ctx, ffi, take_matches, Match, SearchType};

/// reference `src/hamming.rs:390-392` (panics if the lengths differ, `:38, 318`)
pub fn hamming(a: &[u8], b: &[u8]) -> u32 {
    assert!(a.len() == b.len());
    let (ao, bo, mut out) = ([0u64, a.len() as u64], [0u64, b.len() as u64], 0u32);
    check(unsafe { ffi::ta_hamming_batch(ctx(), a.as_ptr(), ao.as_ptr(), b.as_ptr(), bo.as_ptr(), 1, &mut out) });
    out
}

// The scalar, word-wise, movemask and "parallel" variants of the reference return the same value; they are CPU
// micro-variants of one contract (reference `src/hamming.rs:36, 176, 249, 317, 354`).
pub fn hamming_naive(a: &[u8], b: &[u8]) -> u32 { hamming(a, b) }
pub fn hamming_words_64(a: &[u8], b: &[u8]) -> u32 { hamming(a, b) }
pub fn hamming_words_128(a: &[u8], b: &[u8]) -> u32 { hamming(a, b) }
pub fn hamming_simd_parallel(a: &[u8], b: &[u8]) -> u32 { hamming(a, b) }
pub fn hamming_simd_movemask(a: &[u8], b: &[u8]) -> u32 { hamming(a, b) }

/// reference `src/hamming.rs:454-475`.  The reference returns a lazy iterator; the engine materialises the list
/// (same items, same order).  A NUL byte in the haystack panics like `check_no_null_bytes` (`src/lib.rs:237-243`).
pub fn hamming_search_simd_with_opts<'a>(needle: &'a [u8], haystack: &'a [u8], k: u32, search_type: SearchType)
    -> Box<dyn Iterator<Item = Match> + 'a> {
    let off = [0u64, haystack.len() as u64];
    let (mut m, mut mo) = (std::ptr::null_mut(), std::ptr::null_mut());
    let rc = unsafe {
        ffi::ta_hamming_search_batch(ctx(), needle.as_ptr(), needle.len(), haystack.as_ptr(), off.as_ptr(), 1, k,
                                     (search_type == SearchType::Best) as i32, &mut m, &mut mo)
    };
    if rc == ffi::TA_ERR_NUL_BYTE {
        panic!("No zero/null bytes allowed in the string!");
    }
    check(rc);
    Box::new(unsafe { take_matches(m, mo) }.into_iter())
}

/// reference `src/hamming.rs:422-424`: k = needle_len / 2, `SearchType::Best`
pub fn hamming_search_simd<'a>(needle: &'a [u8], haystack: &'a [u8]) -> Box<dyn Iterator<Item = Match> + 'a> {
    hamming_search_simd_with_opts(needle, haystack, unsafe { ffi::ta_search_default_k(needle.len()) }, SearchType::Best)
}

/// reference `src/hamming.rs:588-590`
pub fn hamming_search<'a>(needle: &'a [u8], haystack: &'a [u8]) -> Box<dyn Iterator<Item = Match> + 'a> {
    hamming_search_simd(needle, haystack)
}

/// reference `src/hamming.rs:96-146`: the scalar routine has no NUL-byte restriction (`ta_hamming_search_naive_batch`).
pub fn hamming_search_naive_with_opts<'a>(needle: &'a [u8], haystack: &'a [u8], k: u32, search_type: SearchType)
    -> Box<dyn Iterator<Item = Match> + 'a> {
    let off = [0u64, haystack.len() as u64];
    let (mut m, mut mo) = (std::ptr::null_mut(), std::ptr::null_mut());
    check(unsafe {
        ffi::ta_hamming_search_naive_batch(ctx(), needle.as_ptr(), needle.len(), haystack.as_ptr(), off.as_ptr(), 1, k,
                                           (search_type == SearchType::Best) as i32, &mut m, &mut mo)
    });
    Box::new(unsafe { take_matches(m, mo) }.into_iter())
}

/// reference `src/hamming.rs:70-72`
pub fn hamming_search_naive<'a>(needle: &'a [u8], haystack: &'a [u8]) -> Box<dyn Iterator<Item = Match> + 'a> {
    hamming_search_naive_with_opts(needle, haystack, unsafe { ffi::ta_search_default_k(needle.len()) }, SearchType::Best)
}

/// New: the batch entry point (CSR: bytes + n + 1 offsets per side).
pub fn hamming_batch(a: &[u8], a_off: &[u64], b: &[u8], b_off: &[u64]) -> Vec<u32> {
    assert!(a_off.len() == b_off.len() && !a_off.is_empty());
    let n = a_off.len() - 1;
    let mut out = vec![0u32; n];
    let rc = unsafe {
        ffi::ta_hamming_batch(ctx(), a.as_ptr(), a_off.as_ptr(), b.as_ptr(), b_off.as_ptr(), n, out.as_mut_ptr())
    };
    assert!(rc != ffi::TA_ERR_LEN_MISMATCH, "hamming: a pair differs in length");
    check(rc);
    out
}
