//! reference `src/levenshtein.rs`: every Levenshtein / restricted-Damerau entry point over the batch engine.
use crate::{check, ctx, edit_type, ffi, take_matches, Edit, Match, SearchType};

/// reference `src/levenshtein.rs:20-26`: private fields, `Copy`
#[derive(Copy, Clone, Debug)]
pub struct EditCosts {
    mismatch_cost: u8,
    gap_cost: u8,
    start_gap_cost: u8,
    transpose_cost: Option<u8>,
}

impl EditCosts {
    /// reference `src/levenshtein.rs:38-60` (same asserts)
    pub fn new(mismatch_cost: u8, gap_cost: u8, start_gap_cost: u8, transpose_cost: Option<u8>) -> Self {
        assert!(mismatch_cost > 0);
        assert!(gap_cost > 0);
        if let Some(cost) = transpose_cost {
            assert!(cost > 0);
            assert!((cost >> 1) < mismatch_cost);
            assert!((cost >> 1) < gap_cost);
        }
        Self { mismatch_cost, gap_cost, start_gap_cost, transpose_cost }
    }

    /// reference `src/levenshtein.rs:67-71`
    fn check_search(&self) {
        if let Some(cost) = self.transpose_cost {
            assert!(cost <= self.start_gap_cost + self.gap_cost);
        }
    }

    fn raw(&self) -> ffi::TaCosts {
        ffi::TaCosts {
            mismatch: self.mismatch_cost,
            gap: self.gap_cost,
            start_gap: self.start_gap_cost,
            transpose: self.transpose_cost.unwrap_or(0), // Some(t) requires t > 0, so 0 encodes None
        }
    }
}

/// reference `src/levenshtein.rs:76-81`
pub const LEVENSHTEIN_COSTS: EditCosts =
    EditCosts { mismatch_cost: 1, gap_cost: 1, start_gap_cost: 0, transpose_cost: None };
/// reference `src/levenshtein.rs:84-89`
pub const RDAMERAU_COSTS: EditCosts =
    EditCosts { mismatch_cost: 1, gap_cost: 1, start_gap_cost: 0, transpose_cost: Some(1) };

unsafe fn take_edits(e: *mut ffi::TaEdit, eo: *mut u64) -> Vec<Edit> {
    let n = *eo.add(1) as usize;
    let v = (0..n)
        .map(|i| {
            let x = *e.add(i);
            Edit { edit: edit_type(x.edit), count: x.count as usize }
        })
        .collect();
    ffi::ta_free(e as *mut _);
    ffi::ta_free(eo as *mut _);
    v
}

/// reference `src/levenshtein.rs:714-720`: `None` if the distance exceeds `k`; with `trace_on` the run-length
/// encoded edits of the scalar routine (`:493-606`).
pub fn levenshtein_simd_k_with_opts(a: &[u8], b: &[u8], k: u32, trace_on: bool, costs: EditCosts)
    -> Option<(u32, Option<Vec<Edit>>)> {
    let (ao, bo, mut out) = ([0u64, a.len() as u64], [0u64, b.len() as u64], 0u32);
    if !trace_on {
        check(unsafe {
            ffi::ta_levenshtein_k_batch(ctx(), a.as_ptr(), ao.as_ptr(), b.as_ptr(), bo.as_ptr(), 1, k, costs.raw(), &mut out)
        });
        return if out == ffi::TA_NONE { None } else { Some((out, None)) };
    }
    let (mut e, mut eo) = (std::ptr::null_mut(), std::ptr::null_mut());
    check(unsafe {
        ffi::ta_levenshtein_k_trace_batch(ctx(), a.as_ptr(), ao.as_ptr(), b.as_ptr(), bo.as_ptr(), 1, k, costs.raw(),
                                          &mut out, &mut e, &mut eo)
    });
    let edits = unsafe { take_edits(e, eo) };
    if out == ffi::TA_NONE { None } else { Some((out, Some(edits))) }
}

/// reference `src/levenshtein.rs:677-684`
pub fn levenshtein_simd_k(a: &[u8], b: &[u8], k: u32) -> Option<u32> {
    levenshtein_simd_k_with_opts(a, b, k, false, LEVENSHTEIN_COSTS).map(|r| r.0)
}

/// reference `src/levenshtein.rs:1397-1399`
pub fn levenshtein(a: &[u8], b: &[u8]) -> u32 {
    levenshtein_simd_k(a, b, u32::MAX).unwrap()
}

/// reference `src/levenshtein.rs:1419-1423`
pub fn rdamerau(a: &[u8], b: &[u8]) -> u32 {
    levenshtein_simd_k_with_opts(a, b, u32::MAX, false, RDAMERAU_COSTS).unwrap().0
}

/// reference `src/levenshtein.rs:1480-1494`
pub fn levenshtein_exp_with_opts(a: &[u8], b: &[u8], trace_on: bool, costs: EditCosts) -> (u32, Option<Vec<Edit>>) {
    let (ao, bo, mut out) = ([0u64, a.len() as u64], [0u64, b.len() as u64], 0u32);
    if !trace_on {
        check(unsafe {
            ffi::ta_levenshtein_exp_batch(ctx(), a.as_ptr(), ao.as_ptr(), b.as_ptr(), bo.as_ptr(), 1, costs.raw(), &mut out)
        });
        return (out, None);
    }
    let (mut e, mut eo) = (std::ptr::null_mut(), std::ptr::null_mut());
    check(unsafe {
        ffi::ta_levenshtein_exp_trace_batch(ctx(), a.as_ptr(), ao.as_ptr(), b.as_ptr(), bo.as_ptr(), 1, costs.raw(),
                                            &mut out, &mut e, &mut eo)
    });
    (out, Some(unsafe { take_edits(e, eo) }))
}

/// reference `src/levenshtein.rs:1445-1454`
pub fn levenshtein_exp(a: &[u8], b: &[u8]) -> u32 {
    levenshtein_exp_with_opts(a, b, false, LEVENSHTEIN_COSTS).0
}

/// reference `src/levenshtein.rs:1516-1526`
pub fn rdamerau_exp(a: &[u8], b: &[u8]) -> u32 {
    levenshtein_exp_with_opts(a, b, false, RDAMERAU_COSTS).0
}

/// reference `src/levenshtein.rs:609-624`: chars -> u8 codes in order of first appearance, `None` past 256
fn translate_str(chars: &mut Vec<char>, s: &str) -> Option<Vec<u8>> {
    s.chars()
        .map(|c| match chars.iter().position(|&d| c == d) {
            Some(i) => Some(i as u8),
            None => {
                let idx = chars.len();
                if idx < 256 {
                    chars.push(c);
                    Some(idx as u8)
                } else {
                    None
                }
            }
        })
        .collect()
}

/// reference `src/levenshtein.rs:641-651`
pub fn levenshtein_simd_k_str(a: &str, b: &str, k: u32) -> Option<u32> {
    if a.is_ascii() && b.is_ascii() {
        levenshtein_simd_k(a.as_bytes(), b.as_bytes(), k)
    } else {
        let mut chars = Vec::with_capacity(256);
        let a = translate_str(&mut chars, a)?;
        let b = translate_str(&mut chars, b)?;
        levenshtein_simd_k(&a, &b, k)
    }
}

// ---- the scalar ("naive") names: the same contracts -----------------------------------------------------------------
// The reference's scalar routines are generic over `T: PartialEq`; the engine works on u8, so other element types are
// mapped to u8 codes by equality (at most 256 distinct elements over both inputs).

fn to_codes<T: PartialEq>(a: &[T], b: &[T]) -> (Vec<u8>, Vec<u8>) {
    let mut seen: Vec<&T> = Vec::new();
    fn one<'s, T: PartialEq>(seen: &mut Vec<&'s T>, x: &'s T) -> u8 {
        match seen.iter().position(|d| *d == x) {
            Some(i) => i as u8,
            None => {
                assert!(seen.len() < 256, "more than 256 distinct elements: not representable on the u8 path");
                seen.push(x);
                (seen.len() - 1) as u8
            }
        }
    }
    let ca = a.iter().map(|x| one(&mut seen, x)).collect();
    let cb = b.iter().map(|x| one(&mut seen, x)).collect();
    (ca, cb)
}

/// reference `src/levenshtein.rs:376-607`: the contract itself (the engine is pinned to this routine bit for bit)
pub fn levenshtein_naive_k_with_opts<T: PartialEq>(a: &[T], b: &[T], k: u32, trace_on: bool, costs: EditCosts)
    -> Option<(u32, Option<Vec<Edit>>)> {
    let (ca, cb) = to_codes(a, b);
    levenshtein_simd_k_with_opts(&ca, &cb, k, trace_on, costs)
}

/// reference `src/levenshtein.rs:342-349`
pub fn levenshtein_naive_k(a: &[u8], b: &[u8], k: u32) -> Option<u32> {
    levenshtein_simd_k(a, b, k)
}

/// reference `src/levenshtein.rs:148-319`.  The k = u32::MAX case of the bounded routine, traceback included: the
/// unbounded routine's tie order (a-gap, b-gap if <, substitution if <=, transposition if <=, `:207-249`) is the same
/// decision function as the bounded one's (`:493-532`): substitution wins ties, the a-gap wins ties between gaps.
pub fn levenshtein_naive_with_opts<T: PartialEq>(a: &[T], b: &[T], trace_on: bool, costs: EditCosts)
    -> (u32, Option<Vec<Edit>>) {
    let (ca, cb) = to_codes(a, b);
    levenshtein_simd_k_with_opts(&ca, &cb, u32::MAX, trace_on, costs).unwrap()
}

/// reference `src/levenshtein.rs:105-107`
pub fn levenshtein_naive<T: PartialEq>(a: &[T], b: &[T]) -> u32 {
    levenshtein_naive_with_opts(a, b, false, LEVENSHTEIN_COSTS).0
}

/// reference `src/levenshtein.rs:123-127` (sic)
pub fn levenstein_naive_str(a: &str, b: &str) -> u32 {
    let a: Vec<char> = a.chars().collect();
    let b: Vec<char> = b.chars().collect();
    levenshtein_naive(&a, &b)
}

// ---- search -------------------------------------------------------------------------------------------------------------

/// `SearchType::All` on a long haystack, lazily like the reference's iterator (`src/levenshtein.rs:2282, 2448`;
/// `tests/basic_tests.rs:631-632` takes one `.next()`): the haystack is searched `LAZY_CHUNK` bytes at a time, each
/// chunk restarted `warm = 2 |needle| + start_gap / gap + 2` bytes early -- an optimal alignment of a needle prefix of
/// length j costs at most j gap + start_gap, so it spans at most 2 j + start_gap / gap haystack bytes and every
/// (cost, length) the reference computes for an end position inside the chunk is reproduced by the restarted run (the
/// same argument the engine's own 128-byte work items rest on).  Matches that end in the warm-up belong
/// to the previous chunk and are dropped.
struct LazyAll<'a> {
    needle: &'a [u8],
    haystack: &'a [u8],
    k: u32,
    costs: EditCosts,
    next_chunk: usize, // first end position (exclusive lower bound) of the next chunk
    buf: std::vec::IntoIter<Match>,
}
const LAZY_CHUNK: usize = 1 << 20;

impl<'a> Iterator for LazyAll<'a> {
    type Item = Match;
    fn next(&mut self) -> Option<Match> {
        loop {
            if let Some(m) = self.buf.next() {
                return Some(m);
            }
            if self.next_chunk > self.haystack.len() {
                return None;
            }
            let lo = self.next_chunk; // matches with lo <= end <= hi, except end == 0 which only the first chunk has
            let hi = std::cmp::min(self.haystack.len(), lo + LAZY_CHUNK);
            let raw = self.costs.raw();
            let warm = 2 * self.needle.len() + (raw.start_gap / raw.gap) as usize + 2;
            let from = lo.saturating_sub(warm);
            let part = &self.haystack[from..hi];
            let off = [0u64, part.len() as u64];
            let (mut m, mut mo) = (std::ptr::null_mut(), std::ptr::null_mut());
            check(unsafe {
                ffi::ta_levenshtein_search_batch(ctx(), self.needle.as_ptr(), self.needle.len(), part.as_ptr(),
                                                 off.as_ptr(), 1, self.k, 0, raw, 0, &mut m, &mut mo)
            });
            let first = lo == 0;
            let v: Vec<Match> = unsafe { take_matches(m, mo) }
                .into_iter()
                .filter(|x| if first { true } else { x.end + from > lo })
                .map(|x| Match { start: x.start + from, end: x.end + from, k: x.k })
                .collect();
            self.next_chunk = if hi == self.haystack.len() { hi + 1 } else { hi };
            self.buf = v.into_iter();
        }
    }
}

/// reference `src/levenshtein.rs:1911-1918`.  `All` on an unanchored haystack longer than `LAZY_CHUNK` is lazy like
/// the reference's iterator (see `LazyAll`); everything else is one call whose list is handed out item by item.
pub fn levenshtein_search_simd_with_opts<'a>(needle: &'a [u8], haystack: &'a [u8], k: u32, search_type: SearchType,
                                             costs: EditCosts, anchored: bool) -> Box<dyn Iterator<Item = Match> + 'a> {
    if !needle.is_empty() {
        costs.check_search(); // reference :1965 (after the empty-needle special case)
    }
    if search_type == SearchType::All && !anchored && !needle.is_empty() && haystack.len() > LAZY_CHUNK {
        return Box::new(LazyAll { needle, haystack, k, costs, next_chunk: 0, buf: Vec::new().into_iter() });
    }
    let off = [0u64, haystack.len() as u64];
    let (mut m, mut mo) = (std::ptr::null_mut(), std::ptr::null_mut());
    check(unsafe {
        ffi::ta_levenshtein_search_batch(ctx(), needle.as_ptr(), needle.len(), haystack.as_ptr(), off.as_ptr(), 1, k,
                                         (search_type == SearchType::Best) as i32, costs.raw(), anchored as i32,
                                         &mut m, &mut mo)
    });
    Box::new(unsafe { take_matches(m, mo) }.into_iter())
}

/// reference `src/levenshtein.rs:1866-1878`: k = needle_len / 2, Best, unit costs, unanchored
pub fn levenshtein_search_simd<'a>(needle: &'a [u8], haystack: &'a [u8]) -> Box<dyn Iterator<Item = Match> + 'a> {
    levenshtein_search_simd_with_opts(needle, haystack, unsafe { ffi::ta_search_default_k(needle.len()) },
                                      SearchType::Best, LEVENSHTEIN_COSTS, false)
}

/// reference `src/levenshtein.rs:2508-2513`
pub fn levenshtein_search<'a>(needle: &'a [u8], haystack: &'a [u8]) -> Box<dyn Iterator<Item = Match> + 'a> {
    levenshtein_search_simd(needle, haystack)
}

/// reference `src/levenshtein.rs:1589-1838`: the contract the engine's search is pinned to
pub fn levenshtein_search_naive_with_opts<'a>(needle: &'a [u8], haystack: &'a [u8], k: u32, search_type: SearchType,
                                              costs: EditCosts, anchored: bool) -> Box<dyn Iterator<Item = Match> + 'a> {
    levenshtein_search_simd_with_opts(needle, haystack, k, search_type, costs, anchored)
}

/// reference `src/levenshtein.rs:1549-1556`
pub fn levenshtein_search_naive<'a>(needle: &'a [u8], haystack: &'a [u8]) -> Box<dyn Iterator<Item = Match> + 'a> {
    levenshtein_search_simd(needle, haystack)
}

// ---- new: the batch entry points the GPU is built for (CSR: bytes + n + 1 offsets) ---------------------------------------

pub fn levenshtein_simd_k_batch(a: &[u8], a_off: &[u64], b: &[u8], b_off: &[u64], k: u32, costs: EditCosts)
    -> Vec<Option<u32>> {
    assert!(a_off.len() == b_off.len() && !a_off.is_empty());
    let n = a_off.len() - 1;
    let mut out = vec![0u32; n];
    check(unsafe {
        ffi::ta_levenshtein_k_batch(ctx(), a.as_ptr(), a_off.as_ptr(), b.as_ptr(), b_off.as_ptr(), n, k, costs.raw(),
                                    out.as_mut_ptr())
    });
    out.into_iter().map(|d| if d == ffi::TA_NONE { None } else { Some(d) }).collect()
}

pub fn levenshtein_exp_batch(a: &[u8], a_off: &[u64], b: &[u8], b_off: &[u64], costs: EditCosts) -> Vec<u32> {
    assert!(a_off.len() == b_off.len() && !a_off.is_empty());
    let n = a_off.len() - 1;
    let mut out = vec![0u32; n];
    check(unsafe {
        ffi::ta_levenshtein_exp_batch(ctx(), a.as_ptr(), a_off.as_ptr(), b.as_ptr(), b_off.as_ptr(), n, costs.raw(),
                                      out.as_mut_ptr())
    });
    out
}

/// All matches of one needle in every haystack of a batch: `(matches, offsets)` with the matches of haystack `i` at
/// `matches[offsets[i] .. offsets[i + 1]]`.
pub fn levenshtein_search_batch(needle: &[u8], hay: &[u8], hay_off: &[u64], k: u32, search_type: SearchType,
                                costs: EditCosts, anchored: bool) -> (Vec<Match>, Vec<u64>) {
    assert!(!hay_off.is_empty());
    if !needle.is_empty() {
        costs.check_search();
    }
    let n = hay_off.len() - 1;
    let (mut m, mut mo) = (std::ptr::null_mut(), std::ptr::null_mut());
    check(unsafe {
        ffi::ta_levenshtein_search_batch(ctx(), needle.as_ptr(), needle.len(), hay.as_ptr(), hay_off.as_ptr(), n, k,
                                         (search_type == SearchType::Best) as i32, costs.raw(), anchored as i32,
                                         &mut m, &mut mo)
    });
    unsafe {
        let offs: Vec<u64> = (0..=n).map(|i| *mo.add(i)).collect();
        let total = offs[n] as usize;
        let v = (0..total)
            .map(|i| {
                let x = *m.add(i);
                Match { start: x.start as usize, end: x.end as usize, k: x.k }
            })
            .collect();
        ffi::ta_free(m as *mut _);
        ffi::ta_free(mo as *mut _);
        (v, offs)
    }
}
