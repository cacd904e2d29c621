/*
 * ta_oracle.h -- CPU restatement of triple_accel's SCALAR algorithms (plain C).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The shipped library
 * (libtriple_accel_b200.so) never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED against the reference's own known-answer tests
 * (tests/basic_tests.rs and the doc-tests, transliterated into tests/golden/kat.json).
 * The reference crate itself is Rust-only and cannot be built here (no rustc/cargo in the
 * image, no network), so there is no oracle/_ref binary; see DESIGN.md.
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 */
#ifndef TA_ORACLE_H
#define TA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NONE 0xFFFFFFFFu /* Option::None for a distance */

/* src/levenshtein.rs:20-26.  transpose == 0 means transpose_cost: None. */
typedef struct {
    uint8_t mismatch, gap, start_gap, transpose;
} orc_costs;

/* src/lib.rs:134-142 */
typedef struct {
    uint64_t start, end;
    uint32_t k, _pad;
} orc_match;

/* src/lib.rs:147-165; edit codes follow the declaration order of EditType */
enum { ORC_EDIT_MATCH = 0, ORC_EDIT_MISMATCH = 1, ORC_EDIT_AGAP = 2, ORC_EDIT_BGAP = 3, ORC_EDIT_TRANSPOSE = 4 };
typedef struct {
    uint32_t edit, count;
} orc_edit;

/* EditCosts::new asserts (src/levenshtein.rs:38-60): returns 1 if valid */
int orc_costs_valid(orc_costs c);
/* EditCosts::check_search (src/levenshtein.rs:67-71): returns 1 if valid */
int orc_costs_valid_search(orc_costs c);

/* hamming_naive, src/hamming.rs:36-47.  Returns -1 if lengths differ (the reference panics). */
int64_t orc_hamming_naive(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len);

/* levenshtein_naive_with_opts, src/levenshtein.rs:148-319.
 * If edits_out != NULL a malloc'd RLE traceback is returned (caller frees). */
uint32_t orc_levenshtein_naive_with_opts(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len,
                                         orc_costs c, orc_edit **edits_out, size_t *n_edits_out);

/* levenshtein_naive_k_with_opts, src/levenshtein.rs:376-607.  ORC_NONE when > k. */
uint32_t orc_levenshtein_naive_k_with_opts(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len,
                                           uint32_t k, orc_costs c, orc_edit **edits_out, size_t *n_edits_out);

/* levenshtein_exp_with_opts loop, src/levenshtein.rs:1480-1494 (k = 30, 60, ...) over the scalar k-bounded DP. */
uint32_t orc_levenshtein_exp_with_opts(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len, orc_costs c);

/* levenshtein_search_naive_with_opts, src/levenshtein.rs:1589-1838.
 * search_type: 0 = All, 1 = Best.  Returns the number of matches and a malloc'd array (caller frees),
 * or -1 if costs fail check_search (the reference panics). */
int64_t orc_levenshtein_search_naive_with_opts(const uint8_t *needle, size_t needle_len, const uint8_t *haystack,
                                               size_t haystack_len, uint32_t k, int search_type, orc_costs c,
                                               int anchored, orc_match **out);

/* hamming_search_naive_with_opts, src/hamming.rs:96-146.  Returns the number of matches (malloc'd array). */
int64_t orc_hamming_search_naive_with_opts(const uint8_t *needle, size_t needle_len, const uint8_t *haystack,
                                           size_t haystack_len, uint32_t k, int search_type, orc_match **out);
/* the public entry hamming_search_simd_with_opts, src/hamming.rs:454-475: empty needle -> no matches, NUL byte in
 * the haystack -> -2 (the reference panics, src/lib.rs:237-243), otherwise the scalar routine. */
int64_t orc_hamming_search_with_opts(const uint8_t *needle, size_t needle_len, const uint8_t *haystack,
                                     size_t haystack_len, uint32_t k, int search_type, orc_match **out);
int64_t orc_hamming_search_batch(const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                                 const uint64_t *hay_off, size_t n, uint32_t k, int search_type, orc_match **out,
                                 uint64_t *match_off, int n_threads);

/* default k of levenshtein_search / levenshtein_search_naive, src/levenshtein.rs:1556, 1873 */
uint32_t orc_search_default_k(size_t needle_len);

/* ---- batch drivers (CSR: bytes + n+1 u64 offsets) used by the CPU baseline; pthread parallel-for over pairs ---- */
void orc_hamming_batch(const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off, size_t n,
                       uint32_t *out, int n_threads);
void orc_levenshtein_k_batch(const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                             size_t n, uint32_t k, orc_costs c, uint32_t *out, int n_threads);
void orc_levenshtein_exp_batch(const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                               size_t n, orc_costs c, uint32_t *out, int n_threads);
/* returns total matches; per-haystack counts in counts[n]; matches concatenated into a malloc'd array */
int64_t orc_levenshtein_search_batch(const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                                     const uint64_t *hay_off, size_t n, uint32_t k, int search_type, orc_costs c,
                                     int anchored, orc_match **out, uint64_t *match_off, int n_threads);
void orc_free(void *p);
int orc_max_threads(void);

/* ---- ta_ref_avx2.c: restatement of the reference's AVX2 path (Avx1x32x8 core, Avx::count_mismatches) used as the
 * CPU baseline; NOT the parity oracle (it reproduces the SIMD path's documented deviations from the scalar contract) */
int orc_simd_available(void); /* AVX2 present on this host */
int orc_simd_covers(size_t a_len, size_t b_len, uint32_t k, orc_costs c); /* the reference would pick Avx1x32x8 */
uint32_t orc_levenshtein_simd_k_with_opts(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len, uint32_t k,
                                          orc_costs c, int *covered);
uint32_t orc_levenshtein_simd_exp_with_opts(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len,
                                            orc_costs c);
int64_t orc_hamming_simd(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len);
/* levenshtein_search_simd_with_opts -> the Avx1x32x8 search core (needle <= 32 bytes, u8 cells), else the scalar oracle */
int64_t orc_levenshtein_search_simd_with_opts(const uint8_t *needle, size_t needle_len, const uint8_t *haystack,
                                              size_t haystack_len, uint32_t k, int search_type, orc_costs c,
                                              int anchored, orc_match **out, int *covered);
int64_t orc_levenshtein_search_simd_batch(const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                                          const uint64_t *hay_off, size_t n, uint32_t k, int search_type, orc_costs c,
                                          int anchored, orc_match **out, uint64_t *match_off, int n_threads);
void orc_levenshtein_simd_k_batch(const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                                  size_t n, uint32_t k, orc_costs c, uint32_t *out, int n_threads);
void orc_levenshtein_simd_exp_batch(const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                                    size_t n, orc_costs c, uint32_t *out, int n_threads);
void orc_hamming_simd_batch(const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off, size_t n,
                            uint32_t *out, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
