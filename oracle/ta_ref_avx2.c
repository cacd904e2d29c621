/*
 * ta_ref_avx2.c -- C restatement of the reference's AVX2 (SIMD) code path, for the CPU BASELINE only.
 *
 * THIS IS TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT CODE (same rules as ta_oracle.c).
 *
 * The reference crate is Rust-only and cannot be compiled in this image (no rustc / cargo), so `bench.py --impl
 * reference` cannot time the crate's binary.  What a user of the crate runs on an AVX2 host is not the scalar routine
 * but `levenshtein_simd_k_with_opts` -> `levenshtein_simd_core_avx_1x32x8` (src/levenshtein.rs:714-827, 833-1195)
 * over the `Avx1x32x8` Jewel vector (src/jewel.rs:97-418), and `hamming` -> `Avx::count_mismatches`
 * (src/hamming.rs:317-330, src/jewel.rs:2320-2365).  This file restates exactly those, operation for operation, with
 * the same intrinsics, so that the baseline beside the GPU number is the reference's fast path and not its fallback.
 * Labelled everywhere as "restatement, not the reference binary" (cpu_baseline.kind = "port").
 *
 * Scope: the single-register u8 type `Avx1x32x8`, which every BASELINE.json configuration selects (SURVEY.md
 * Appendix A: unit_k <= 30 and max_k <= 254).  Wider bands (the reference's Avx2x32x8 ... AvxNx8x32 types) fall back
 * to the scalar oracle here and the caller is told through orc_simd_covers().  No traceback (trace_on = false).
 *
 * Known, documented difference from the scalar contract (hard part 2): with a transposition cost the
 * SIMD core blends the transposition candidate in unconditionally (src/levenshtein.rs:1093-1097), so on small
 * alphabets it can return a larger distance than the scalar routine.  The GPU path follows the scalar contract; this
 * file follows the SIMD code, because its job is to cost what the reference costs.
 */
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ta_oracle.h"

#define TGT __attribute__((target("avx2")))

int orc_simd_available(void) { return __builtin_cpu_supports("avx2") ? 1 : 0; }

/* max_k / unit_k of levenshtein_simd_k_with_opts, src/levenshtein.rs:729-763 */
static void simd_bounds(size_t a_len, size_t b_len, uint32_t k, orc_costs c, uint32_t *max_k_out, uint32_t *unit_k_out) {
    const uint32_t min_len = (uint32_t)(a_len < b_len ? a_len : b_len);
    const uint32_t max_len = (uint32_t)(a_len < b_len ? b_len : a_len);
    uint32_t gaps = (min_len << 1) * c.gap;
    if (min_len != 0) gaps += c.start_gap + (max_len == min_len ? c.start_gap : 0u);
    uint32_t max_k = min_len * c.mismatch < gaps ? min_len * c.mismatch : gaps;
    const uint32_t bound = max_k + (max_len - min_len) * c.gap + (max_len == min_len ? 0u : c.start_gap);
    max_k = k < bound ? k : bound;
    uint32_t unit_k = (max_k > c.start_gap ? max_k - c.start_gap : 0u) / c.gap;
    if (unit_k > max_len) unit_k = max_len;
    *max_k_out = max_k;
    *unit_k_out = unit_k;
}

/* 1 if the single-register core handles (lengths, k, costs), i.e. the reference would pick Avx1x32x8
 * (src/levenshtein.rs:766-771: unit_k <= 32 - 2 and max_k <= 254) */
int orc_simd_covers(size_t a_len, size_t b_len, uint32_t k, orc_costs c) {
    uint32_t max_k, unit_k;
    if (a_len == 0 && b_len == 0) return 1;
    simd_bounds(a_len, b_len, k, c, &max_k, &unit_k);
    return unit_k <= 30 && max_k <= 254;
}

/* Jewel shifts for one __m256i (src/jewel.rs:245-291 with a single register) */
TGT static inline __m256i shl1(__m256i v) { /* lane i <- lane i + 1, zero into lane 31 */
    return _mm256_alignr_epi8(_mm256_permute2x128_si256(v, v, 0x81), v, 1);
}
TGT static inline __m256i shr1(__m256i v) { /* lane i <- lane i - 1, zero into lane 0 */
    return _mm256_alignr_epi8(v, _mm256_permute2x128_si256(v, v, 0x08), 15);
}
/* slow_insert / slow_loadu go through a stack array in the reference as well (src/jewel.rs:138-185) */
TGT static inline __m256i slow_insert(__m256i v, int i, uint32_t val) {
    uint8_t arr[32];
    _mm256_storeu_si256((__m256i *)arr, v);
    arr[i] = (uint8_t)val;
    return _mm256_loadu_si256((const __m256i *)arr);
}
TGT static inline __m256i slow_loadu(__m256i v, int idx, const uint8_t *p, size_t len, int reverse) {
    uint8_t arr[32];
    if (len == 0) return v;
    _mm256_storeu_si256((__m256i *)arr, v);
    for (size_t i = 0; i < len; i++) arr[reverse ? idx - (int)i : idx + (int)i] = p[i];
    return _mm256_loadu_si256((const __m256i *)arr);
}

/* levenshtein_simd_core_avx_1x32x8 without traceback, src/levenshtein.rs:841-1168; `k` is the caller's max_k */
TGT static uint32_t simd_core_1x32x8(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len, uint32_t k,
                                     orc_costs costs) {
    if (a_len > b_len) { /* :852-853 */
        const uint8_t *tp = a;
        a = b;
        b = tp;
        size_t tl = a_len;
        a_len = b_len;
        b_len = tl;
    }
    size_t unit_k = (k > costs.start_gap ? k - costs.start_gap : 0u) / costs.gap; /* :856-859 */
    if (unit_k > b_len) unit_k = b_len;
    if (b_len - a_len > unit_k) return ORC_NONE; /* :861-863 */

    const __m256i ones = _mm256_set1_epi8(-1);
    __m256i dp1 = ones, dp2 = ones, dp0 = ones, dp_temp = ones; /* :867-871 */
    __m256i a_gap_dp = ones, b_gap_dp = ones;                   /* :875-876 */
    const int k1_div2 = 15, k2_div2 = 15;                       /* max_len = 32: k1 = 31, k2 = 30 (:880-883) */
    const uint32_t open = (uint32_t)costs.start_gap + costs.gap;
    dp1 = slow_insert(dp1, k1_div2, 0); /* :886-898 */
    dp2 = slow_insert(dp2, k2_div2 - 1, open);
    dp2 = slow_insert(dp2, k2_div2, open);
    b_gap_dp = slow_insert(b_gap_dp, k2_div2 - 1, open);
    a_gap_dp = slow_insert(a_gap_dp, k2_div2, open);

    const __m256i zeros = _mm256_setzero_si256();
    __m256i a_k1 = slow_loadu(zeros, k1_div2 - 1, a, a_len < 15 ? a_len : 15, 1); /* :904-914 */
    __m256i b_k1 = slow_loadu(zeros, k1_div2 + 1, b, b_len < 15 ? b_len : 15, 0);
    __m256i a_k2 = slow_loadu(zeros, k2_div2 - 1, a, a_len < 15 ? a_len : 15, 1);
    __m256i b_k2 = slow_loadu(zeros, k2_div2, b, b_len < 15 ? b_len : 15, 0);
    size_t k1_idx = k1_div2 - 1, k2_idx = k2_div2 - 1; /* :917-918 */

    const size_t len_diff = b_len - a_len; /* :920-933 */
    const size_t len = a_len + b_len + 1;
    const size_t len_div2 = (len >> 1) + (len & 1);
    const int ends_with_k2 = (len & 1) == 0;
    const size_t final_idx = ends_with_k2 ? k2_div2 + ((len_diff - 1) >> 1) : k1_div2 + (len_diff >> 1);

    __m256i match_mask0 = zeros, match_mask1, sub, a_gap, b_gap, transpose = zeros; /* :955-960 */
    const __m256i mismatch_cost = _mm256_set1_epi8((char)costs.mismatch);           /* :962-969 */
    const __m256i gap_cost = _mm256_set1_epi8((char)costs.gap);
    const __m256i start_gap_cost = _mm256_set1_epi8((char)open);
    const __m256i transpose_cost = _mm256_set1_epi8((char)costs.transpose);
    const int allow_transpose = costs.transpose != 0;

    for (size_t it = 1; it < len_div2; it++) { /* :1022 */
        k1_idx++;
        k2_idx++;
        a_k1 = shr1(a_k1); /* :1027-1049: move the windows */
        if (k1_idx < a_len) a_k1 = _mm256_insert_epi8(a_k1, (char)a[k1_idx], 0);
        b_k1 = shl1(b_k1);
        if (k1_idx < b_len) b_k1 = _mm256_insert_epi8(b_k1, (char)b[k1_idx], 30);
        a_k2 = shr1(a_k2);
        if (k2_idx < a_len) a_k2 = _mm256_insert_epi8(a_k2, (char)a[k2_idx], 0);
        b_k2 = shl1(b_k2);
        if (k2_idx < b_len) b_k2 = _mm256_insert_epi8(b_k2, (char)b[k2_idx], 29);

        /* k1 diagonal, :1052-1103 */
        match_mask1 = _mm256_cmpeq_epi8(a_k1, b_k1);
        sub = _mm256_andnot_si256(match_mask1, mismatch_cost);
        sub = _mm256_adds_epu8(sub, dp1);
        a_gap = _mm256_adds_epu8(dp2, start_gap_cost);
        a_gap_dp = _mm256_adds_epu8(a_gap_dp, gap_cost);
        a_gap_dp = _mm256_min_epu8(a_gap_dp, a_gap);
        a_gap_dp = shr1(a_gap_dp);
        a_gap_dp = _mm256_insert_epi8(a_gap_dp, -1, 0);
        b_gap = _mm256_adds_epu8(dp2, start_gap_cost);
        b_gap_dp = _mm256_adds_epu8(b_gap_dp, gap_cost);
        b_gap_dp = _mm256_min_epu8(b_gap_dp, b_gap);
        if (allow_transpose) {
            transpose = shr1(match_mask0);
            transpose = _mm256_and_si256(transpose, match_mask0);
            match_mask0 = _mm256_andnot_si256(match_mask1, transpose);
            transpose = _mm256_adds_epu8(dp0, transpose_cost);
        }
        dp0 = _mm256_min_epu8(a_gap_dp, b_gap_dp);
        dp0 = _mm256_min_epu8(dp0, sub);
        if (allow_transpose) {
            dp0 = _mm256_blendv_epi8(dp0, transpose, match_mask0);
            __m256i t = match_mask0;
            match_mask0 = match_mask1;
            match_mask1 = t;
        }
        { /* :1105-1107: dp0 -> dp_temp -> dp1 -> dp2 rotate */
            __m256i t = dp0;
            dp0 = dp_temp;
            dp_temp = dp1;
            dp1 = dp2;
            dp2 = t;
        }

        /* k2 diagonal, :1110-1161 */
        match_mask1 = _mm256_cmpeq_epi8(a_k2, b_k2);
        sub = _mm256_andnot_si256(match_mask1, mismatch_cost);
        sub = _mm256_adds_epu8(sub, dp1);
        b_gap = _mm256_adds_epu8(dp2, start_gap_cost);
        b_gap_dp = _mm256_adds_epu8(b_gap_dp, gap_cost);
        b_gap_dp = _mm256_min_epu8(b_gap_dp, b_gap);
        b_gap_dp = shl1(b_gap_dp);
        b_gap_dp = _mm256_insert_epi8(b_gap_dp, -1, 31);
        a_gap = _mm256_adds_epu8(dp2, start_gap_cost);
        a_gap_dp = _mm256_adds_epu8(a_gap_dp, gap_cost);
        a_gap_dp = _mm256_min_epu8(a_gap_dp, a_gap);
        if (allow_transpose) {
            transpose = shl1(match_mask0);
            transpose = _mm256_and_si256(transpose, match_mask0);
            match_mask0 = _mm256_andnot_si256(match_mask1, transpose);
            transpose = _mm256_adds_epu8(dp0, transpose_cost);
        }
        dp0 = _mm256_min_epu8(a_gap_dp, b_gap_dp);
        dp0 = _mm256_min_epu8(dp0, sub);
        if (allow_transpose) {
            dp0 = _mm256_blendv_epi8(dp0, transpose, match_mask0);
            __m256i t = match_mask0;
            match_mask0 = match_mask1;
            match_mask1 = t;
        }
        {
            __m256i t = dp0;
            dp0 = dp_temp;
            dp_temp = dp1;
            dp1 = dp2;
            dp2 = t;
        }
    }
    uint8_t arr[32]; /* slow_extract, :1164-1168 */
    _mm256_storeu_si256((__m256i *)arr, ends_with_k2 ? dp2 : dp1);
    const uint32_t final_res = arr[final_idx];
    return final_res > k ? ORC_NONE : final_res;
}

/* levenshtein_simd_k_with_opts (trace_on = false), src/levenshtein.rs:714-827.  *covered = 0 when the reference would
 * have used a wider Jewel type than the one restated here (the scalar oracle answers instead). */
uint32_t orc_levenshtein_simd_k_with_opts(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len, uint32_t k,
                                          orc_costs c, int *covered) {
    if (covered) *covered = 1;
    if (a_len == 0 && b_len == 0) return 0;
    uint32_t max_k, unit_k;
    simd_bounds(a_len, b_len, k, c, &max_k, &unit_k);
    if (orc_simd_available() && unit_k <= 30 && max_k <= 254) return simd_core_1x32x8(a, a_len, b, b_len, max_k, c);
    if (covered) *covered = 0;
    return orc_levenshtein_naive_k_with_opts(a, a_len, b, b_len, k, c, NULL, NULL);
}

/* levenshtein_exp_with_opts / rdamerau_exp over the SIMD k routine, src/levenshtein.rs:1480-1494, 1516-1526 */
uint32_t orc_levenshtein_simd_exp_with_opts(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len,
                                            orc_costs c) {
    uint32_t k = 30;
    for (;;) {
        const uint32_t r = orc_levenshtein_simd_k_with_opts(a, a_len, b, b_len, k, c, NULL);
        if (r != ORC_NONE) return r;
        k *= 2;
    }
}

/* Avx::count_mismatches, src/jewel.rs:2320-2365 (called by hamming_simd_parallel, src/hamming.rs:317-330) */
TGT static uint32_t count_mismatches_avx2(const uint8_t *a, const uint8_t *b, size_t len) {
    const size_t refresh_len = len / (255 * 32);
    const __m256i zeros = _mm256_setzero_si256();
    __m256i sad = zeros;
    const __m256i *pa = (const __m256i *)a, *pb = (const __m256i *)b;
    for (size_t i = 0; i < refresh_len; i++) {
        __m256i curr = zeros;
        for (size_t j = i * 255; j < (i + 1) * 255; j++)
            curr = _mm256_sub_epi8(curr, _mm256_cmpeq_epi8(_mm256_loadu_si256(pa + j), _mm256_loadu_si256(pb + j)));
        sad = _mm256_add_epi64(sad, _mm256_sad_epu8(curr, zeros));
    }
    const size_t word_len = len >> 5;
    __m256i curr = zeros;
    for (size_t i = refresh_len * 255; i < word_len; i++)
        curr = _mm256_sub_epi8(curr, _mm256_cmpeq_epi8(_mm256_loadu_si256(pa + i), _mm256_loadu_si256(pb + i)));
    sad = _mm256_add_epi64(sad, _mm256_sad_epu8(curr, zeros));
    uint32_t sad_arr[8];
    _mm256_storeu_si256((__m256i *)sad_arr, sad);
    uint32_t res = sad_arr[0] + sad_arr[2] + sad_arr[4] + sad_arr[6];
    for (size_t i = word_len << 5; i < len; i++) res += a[i] == b[i];
    return (uint32_t)len - res;
}

/* hamming_simd_parallel; -1 if the lengths differ (the reference panics) */
int64_t orc_hamming_simd(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len) {
    if (a_len != b_len) return -1;
    if (orc_simd_available()) return (int64_t)count_mismatches_avx2(a, b, a_len);
    return orc_hamming_naive(a, a_len, b, b_len);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* levenshtein_search_simd_with_opts -> levenshtein_search_simd_core_avx_1x32x8, src/levenshtein.rs:1911-2155,  */
/* 2161-2449, with Jewel's double_min_length / triple_min_length (src/jewel.rs:362-417).  Restated for the      */
/* single-register u8 type only (needle <= 32 bytes and max(needle_len + unit_k, k + 1) <= 255: every search    */
/* configuration in BASELINE.json); anything else falls back to the scalar oracle (*covered = 0).  The SIMD     */
/* path's length tie-break differs from the scalar routine's in rare cases: this is a timing  */
/* baseline, not the parity oracle.                                                                             */

TGT static inline __m256i shl2(__m256i v) { /* lane i <- lane i + 2, zeros into lanes 30, 31 */
    return _mm256_alignr_epi8(_mm256_permute2x128_si256(v, v, 0x81), v, 2);
}
TGT static inline void double_min_length(__m256i new_gap, __m256i *cont_gap, __m256i new_len, __m256i *cont_len) {
    const __m256i res_min = _mm256_min_epu8(new_gap, *cont_gap);
    const __m256i new_mask = _mm256_cmpeq_epi8(new_gap, res_min);
    __m256i res_len = _mm256_blendv_epi8(*cont_len, new_len, new_mask);
    const __m256i eq_mask = _mm256_cmpeq_epi8(new_gap, *cont_gap);
    res_len = _mm256_blendv_epi8(res_len, _mm256_max_epu8(new_len, *cont_len), eq_mask);
    *cont_gap = res_min;
    *cont_len = res_len;
}
TGT static inline void triple_min_length(__m256i sub, __m256i a_gap, __m256i b_gap, __m256i sub_len, __m256i a_len,
                                         __m256i b_len, __m256i *res_min, __m256i *res_len) {
    const __m256i min1 = _mm256_min_epu8(a_gap, b_gap);
    const __m256i a_mask = _mm256_cmpeq_epi8(a_gap, min1);
    __m256i len1 = _mm256_blendv_epi8(b_len, a_len, a_mask);
    len1 = _mm256_blendv_epi8(len1, _mm256_max_epu8(a_len, b_len), _mm256_cmpeq_epi8(a_gap, b_gap));
    const __m256i min2 = _mm256_min_epu8(sub, min1);
    __m256i len2 = _mm256_blendv_epi8(len1, sub_len, _mm256_cmpeq_epi8(sub, min2));
    len2 = _mm256_blendv_epi8(len2, _mm256_max_epu8(sub_len, len1), _mm256_cmpeq_epi8(sub, min1));
    *res_min = min2;
    *res_len = len2;
}

typedef struct {
    orc_match *v;
    size_t n, cap;
} match_vec;
static void mv_push(match_vec *m, uint64_t start, uint64_t end, uint32_t k) {
    if (m->n == m->cap) {
        m->cap = m->cap ? 2 * m->cap : 16;
        m->v = (orc_match *)realloc(m->v, m->cap * sizeof(orc_match));
    }
    m->v[m->n].start = start, m->v[m->n].end = end, m->v[m->n].k = k, m->v[m->n]._pad = 0;
    m->n++;
}

TGT static int64_t search_core_1x32x8(const uint8_t *needle, size_t needle_len, const uint8_t *haystack,
                                      size_t haystack_len, uint32_t k, int best, orc_costs costs, int anchored,
                                      orc_match **out) {
    const __m256i ones8 = _mm256_set1_epi8(-1), zeros = _mm256_setzero_si256();
    __m256i dp0 = ones8, dp_temp = ones8, dp1 = ones8, dp2 = ones8, needle_gap_dp = ones8, haystack_gap_dp = ones8;
    const uint32_t open = (uint32_t)costs.start_gap + costs.gap;
    dp2 = slow_insert(dp2, 31, open); /* :2185-2192 */
    haystack_gap_dp = slow_insert(haystack_gap_dp, 31, open);
    __m256i length0 = zeros, length_temp = zeros, length1 = zeros, length2 = zeros, needle_gap_length = zeros,
            haystack_gap_length = zeros;
    const __m256i ones = _mm256_set1_epi8(1), twos = _mm256_set1_epi8(2);
    size_t len; /* :2206-2218 */
    if (anchored) {
        size_t extra = (size_t)((k > costs.start_gap ? k - costs.start_gap : 0u) / costs.gap);
        size_t lim = needle_len + extra < needle_len ? (size_t)-1 : needle_len + extra;
        len = needle_len + (haystack_len < lim ? haystack_len : lim);
    } else {
        len = needle_len + haystack_len;
    }
    const int final_idx = 32 - (int)needle_len;
    __m256i needle_window = slow_loadu(zeros, 31, needle, needle_len, 1); /* :2223-2229 */
    __m256i haystack_window = zeros;
    size_t haystack_idx = 0;
    uint32_t curr_k = k;
    __m256i match_mask0 = zeros, match_mask1, match_mask_cost, sub, sub_length, needle_gap, haystack_gap,
            transpose = zeros, transpose_length = zeros;
    const __m256i mismatch_cost = _mm256_set1_epi8((char)costs.mismatch), gap_cost = _mm256_set1_epi8((char)costs.gap),
                  start_gap_cost = _mm256_set1_epi8((char)costs.start_gap),
                  transpose_cost = _mm256_set1_epi8((char)costs.transpose);
    const int allow_transpose = costs.transpose != 0;
    match_vec res = {0, 0, 0};
#define CAPK(x) ((x) < k + 1 ? (x) : k + 1)
    for (size_t i = 1; i < len;) { /* :2283 */
        haystack_window = shl1(haystack_window);
        if (haystack_idx < haystack_len) {
            haystack_window = _mm256_insert_epi8(haystack_window, (char)haystack[haystack_idx], 31); /* insert_last_0 */
            haystack_idx++;
        }
        match_mask1 = _mm256_cmpeq_epi8(needle_window, haystack_window);
        match_mask_cost = _mm256_andnot_si256(match_mask1, mismatch_cost);
        sub = shl1(dp1); /* match/mismatch, :2297-2311 */
        if (anchored && i > 1) sub = _mm256_insert_epi8(sub, (char)CAPK((uint32_t)(i - 1) * costs.gap + costs.start_gap), 31);
        sub = _mm256_adds_epu8(sub, match_mask_cost);
        sub_length = _mm256_add_epi8(shl1(length1), ones);
        needle_gap = _mm256_adds_epu8(dp2, start_gap_cost); /* gap in needle, :2314-2322 */
        double_min_length(needle_gap, &needle_gap_dp, length2, &needle_gap_length);
        needle_gap_dp = _mm256_adds_epu8(needle_gap_dp, gap_cost);
        needle_gap_length = _mm256_add_epi8(needle_gap_length, ones);
        haystack_gap = _mm256_adds_epu8(dp2, start_gap_cost); /* gap in haystack, :2325-2346 */
        double_min_length(haystack_gap, &haystack_gap_dp, length2, &haystack_gap_length);
        haystack_gap_dp = shl1(haystack_gap_dp);
        if (anchored)
            haystack_gap_dp = _mm256_insert_epi8(haystack_gap_dp, (char)CAPK((uint32_t)i * costs.gap + costs.start_gap), 31);
        else
            haystack_gap_dp = _mm256_insert_epi8(haystack_gap_dp, (char)costs.start_gap, 31);
        haystack_gap_dp = _mm256_adds_epu8(haystack_gap_dp, gap_cost);
        haystack_gap_length = shl1(haystack_gap_length);
        if (allow_transpose) { /* :2348-2367 */
            transpose = _mm256_and_si256(shl1(match_mask0), match_mask0);
            match_mask0 = _mm256_andnot_si256(match_mask1, transpose);
            dp0 = shl2(dp0);
            if (anchored && i > 3) dp0 = _mm256_insert_epi8(dp0, (char)CAPK((uint32_t)(i - 3) * costs.gap + costs.start_gap), 30);
            length0 = shl2(length0);
            transpose = _mm256_adds_epu8(dp0, transpose_cost);
            transpose_length = _mm256_add_epi8(length0, twos);
        }
        triple_min_length(sub, needle_gap_dp, haystack_gap_dp, sub_length, needle_gap_length, haystack_gap_length, &dp0,
                          &length0);
        if (allow_transpose) {
            dp0 = _mm256_blendv_epi8(dp0, transpose, match_mask0);
            length0 = _mm256_blendv_epi8(length0, transpose_length, match_mask0);
            __m256i t = match_mask0;
            match_mask0 = match_mask1;
            match_mask1 = t;
        }
        { /* :2388-2393 */
            __m256i t = dp0;
            dp0 = dp_temp, dp_temp = dp1, dp1 = dp2, dp2 = t;
            t = length0;
            length0 = length_temp, length_temp = length1, length1 = length2, length2 = t;
        }
        i++;
        if (i >= needle_len) { /* :2397-2420 */
            uint8_t a1[32], a2[32];
            _mm256_storeu_si256((__m256i *)a1, dp2);
            _mm256_storeu_si256((__m256i *)a2, length2);
            const uint32_t final_res = a1[final_idx];
            const size_t final_length = a2[final_idx];
            if (final_res <= curr_k) {
                const size_t end_idx = i - needle_len;
                if (best) curr_k = final_res;
                if (!best || res.n == 0 || end_idx - final_length > res.v[res.n - 1].start)
                    mv_push(&res, end_idx - final_length, end_idx, final_res); /* :2429-2441 folded in */
                else
                    res.v[res.n - 1].start = end_idx - final_length, res.v[res.n - 1].end = end_idx,
                                   res.v[res.n - 1].k = final_res;
            }
        }
    }
#undef CAPK
    if (best) { /* only retain matches with the lowest k, :2447 */
        size_t w = 0;
        for (size_t r = 0; r < res.n; r++)
            if (res.v[r].k == curr_k) res.v[w++] = res.v[r];
        res.n = w;
    }
    *out = res.v;
    return (int64_t)res.n;
}

/* public entry; *covered = 0 when the scalar oracle answered instead */
int64_t orc_levenshtein_search_simd_with_opts(const uint8_t *needle, size_t needle_len, const uint8_t *haystack,
                                              size_t haystack_len, uint32_t k, int search_type, orc_costs c,
                                              int anchored, orc_match **out, int *covered) {
    if (covered) *covered = 0;
    if (needle_len != 0 && orc_costs_valid_search(c) && orc_simd_available()) {
        const uint32_t unit_k = (k > c.start_gap ? k - c.start_gap : 0u) / c.gap;
        uint64_t ub1 = (uint64_t)needle_len + unit_k, ub2 = (uint64_t)k + 1;
        const uint64_t upper = ub1 > ub2 ? ub1 : ub2;
        if (needle_len <= 32 && upper <= 255) {
            if (covered) *covered = 1;
            return search_core_1x32x8(needle, needle_len, haystack, haystack_len, k, search_type == 1, c, anchored, out);
        }
    }
    return orc_levenshtein_search_naive_with_opts(needle, needle_len, haystack, haystack_len, k, search_type, c,
                                                  anchored, out);
}
