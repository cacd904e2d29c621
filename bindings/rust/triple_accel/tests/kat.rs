// A few of the reference's own known-answer tests (tests/basic_tests.rs and the doc-tests), run against the shim on a
// box with a Rust toolchain, a B200 and libtriple_accel_b200.so:
//     TRIPLE_ACCEL_B200_LIB_DIR=<repo>/triple_accel_b200 cargo test
use triple_accel::levenshtein::*;
use triple_accel::*;

#[test]
fn hamming_kats() {
    assert_eq!(hamming(b"abc", b"abd"), 1);
    assert_eq!(hamming_naive(b"abc", b"abd"), 1);
    assert_eq!(hamming_simd_parallel(b"", b""), 0);
}

#[test]
fn levenshtein_kats() {
    assert_eq!(levenshtein(b"abc", b"bcd"), 2);
    assert_eq!(levenshtein_exp(b"abc", b"bcd"), 2);
    assert_eq!(rdamerau(b"abc", b"acb"), 1);
    assert_eq!(levenshtein_simd_k(b"abc", b"ab", 1), Some(1));
    assert_eq!(levenshtein_simd_k(b"abc", b"xyz", 2), None);
    let res = levenshtein_simd_k_with_opts(b"abc", b"acb", 2, true, RDAMERAU_COSTS).unwrap();
    assert_eq!(res.0, 1);
    assert_eq!(
        res.1.unwrap(),
        vec![Edit { edit: EditType::Match, count: 1 }, Edit { edit: EditType::Transpose, count: 1 }]
    );
    assert_eq!(levenshtein_simd_k_str("abc", "ab", 1), Some(1));
    assert_eq!(levenstein_naive_str("abc", "ab"), 1);
}

#[test]
fn search_kats() {
    let matches: Vec<Match> = levenshtein_search(b"abc", b"  abd").collect();
    assert_eq!(matches, vec![Match { start: 2, end: 5, k: 1 }]);
    let matches: Vec<Match> = hamming_search(b"abc", b"  abd").collect();
    assert_eq!(matches, vec![Match { start: 2, end: 5, k: 1 }]);
}

#[test]
fn batch_entry_points() {
    let (a, ao) = (b"abcabd".to_vec(), vec![0u64, 3, 6]);
    let (b, bo) = (b"abdxyz".to_vec(), vec![0u64, 3, 6]);
    assert_eq!(hamming_batch(&a, &ao, &b, &bo), vec![1, 3]);
    assert_eq!(levenshtein_simd_k_batch(&a, &ao, &b, &bo, 2, LEVENSHTEIN_COSTS), vec![Some(1), None]);
    assert_eq!(levenshtein_exp_batch(&a, &ao, &b, &bo, LEVENSHTEIN_COSTS), vec![1, 3]);
}
