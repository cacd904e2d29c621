/*
 * ta_oracle.c -- CPU restatement of triple_accel's scalar algorithms.  TEST INFRASTRUCTURE ONLY
 * (see ta_oracle.h).  Arithmetic is u32 with the same wrap/saturate behaviour as the Rust release
 * build; every function names the reference lines it restates.
 */
#include "ta_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define U32_MAX 0xFFFFFFFFu

static inline uint32_t sat_add(uint32_t x, uint32_t y) { /* u32::saturating_add */
    uint32_t s = x + y;
    return s < x ? U32_MAX : s;
}
static inline uint32_t sat_sub(uint32_t x, uint32_t y) { return x > y ? x - y : 0; }
static inline uint32_t min_u32(uint32_t x, uint32_t y) { return x < y ? x : y; }
static inline size_t min_sz(size_t x, size_t y) { return x < y ? x : y; }
static inline size_t max_sz(size_t x, size_t y) { return x > y ? x : y; }

int orc_costs_valid(orc_costs c) { /* src/levenshtein.rs:44-52 */
    if (c.mismatch == 0 || c.gap == 0) return 0;
    if (c.transpose != 0) {
        if ((c.transpose >> 1) >= c.mismatch) return 0;
        if ((c.transpose >> 1) >= c.gap) return 0;
    }
    return 1;
}

int orc_costs_valid_search(orc_costs c) { /* src/levenshtein.rs:67-71 (u8 + u8 as the reference writes it) */
    if (c.transpose != 0) {
        uint32_t lim = (uint32_t)c.start_gap + (uint32_t)c.gap;
        if (lim > 255) return 0; /* u8 overflow: debug panic; release would wrap.  Treat as invalid. */
        if (c.transpose > lim) return 0;
    }
    return 1;
}

int64_t orc_hamming_naive(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len) {
    /* src/hamming.rs:36-47 */
    if (a_len != b_len) return -1;
    uint32_t res = 0;
    for (size_t i = 0; i < a_len; i++) res += (a[i] != b[i]);
    return (int64_t)res;
}

/* RLE push used by both tracebacks (src/levenshtein.rs:307-311, 598-602) */
typedef struct {
    orc_edit *v;
    size_t n, cap;
} edit_vec;
static void ev_push(edit_vec *e, uint32_t edit) {
    if (e->n > 0 && e->v[e->n - 1].edit == edit) {
        e->v[e->n - 1].count += 1;
        return;
    }
    if (e->n == e->cap) {
        e->cap = e->cap ? e->cap * 2 : 16;
        e->v = (orc_edit *)realloc(e->v, e->cap * sizeof(orc_edit));
    }
    e->v[e->n].edit = edit;
    e->v[e->n].count = 1;
    e->n++;
}
static void ev_reverse(edit_vec *e) {
    for (size_t i = 0, j = e->n; i + 1 < j; i++) {
        j--;
        orc_edit t = e->v[i];
        e->v[i] = e->v[j];
        e->v[j] = t;
    }
}

uint32_t orc_levenshtein_naive_with_opts(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len,
                                         orc_costs c, orc_edit **edits_out, size_t *n_edits_out) {
    /* src/levenshtein.rs:157-319 */
    int trace_on = edits_out != NULL;
    int swap = a_len > b_len; /* :157 */
    const uint8_t *an = swap ? b : a;
    size_t an_len = swap ? b_len : a_len;
    const uint8_t *bn = swap ? a : b;
    size_t bn_len = swap ? a_len : b_len;
    uint32_t mism = c.mismatch, gap = c.gap, sgap = c.start_gap, tcost = c.transpose;
    int allow_t = c.transpose != 0;

    size_t len = an_len + 1;
    uint32_t *dp0 = (uint32_t *)calloc(len, 4), *dp1 = (uint32_t *)calloc(len, 4), *dp2 = (uint32_t *)calloc(len, 4);
    uint32_t *a_gap = (uint32_t *)malloc(len * 4), *b_gap = (uint32_t *)malloc(len * 4);
    for (size_t i = 0; i < len; i++) a_gap[i] = b_gap[i] = U32_MAX; /* :175-176 */
    uint8_t *tb = trace_on ? (uint8_t *)calloc((bn_len + 1) * len, 1) : NULL;

    for (size_t i = 0; i < len; i++) { /* :183-189 */
        dp1[i] = (uint32_t)i * gap + (i == 0 ? 0 : sgap);
        if (trace_on) tb[i] = 2;
    }

    for (size_t i = 1; i < bn_len + 1; i++) { /* :191-253 */
        a_gap[0] = (uint32_t)i * gap + sgap;
        dp2[0] = (uint32_t)i * gap + sgap;
        if (trace_on) tb[i * len] = 1;

        for (size_t j = 1; j < len; j++) {
            uint32_t sub = dp1[j - 1] + (uint32_t)(an[j - 1] != bn[i - 1]) * mism;
            a_gap[j] = min_u32(dp1[j] + sgap + gap, sat_add(a_gap[j], gap));
            b_gap[j] = min_u32(dp2[j - 1] + sgap + gap, sat_add(b_gap[j - 1], gap));
            size_t ti = i * len + j;

            dp2[j] = a_gap[j];
            if (trace_on) tb[ti] = 1;

            if (b_gap[j] < dp2[j]) {
                dp2[j] = b_gap[j];
                if (trace_on) tb[ti] = 2;
            }
            if (sub <= dp2[j]) {
                dp2[j] = sub;
                if (trace_on) tb[ti] = 0;
            }
            if (allow_t && i > 1 && j > 1 && an[j - 1] == bn[i - 2] && an[j - 2] == bn[i - 1]) {
                uint32_t tr = dp0[j - 2] + tcost;
                if (tr <= dp2[j]) {
                    dp2[j] = tr;
                    if (trace_on) tb[ti] = 3;
                }
            }
        }
        uint32_t *t = dp0; /* :251-252 */
        dp0 = dp1;
        dp1 = dp2;
        dp2 = t;
    }

    uint32_t res = dp1[an_len];

    if (trace_on) { /* :255-315 */
        edit_vec ev = {NULL, 0, 0};
        size_t i = bn_len, j = an_len;
        while (i > 0 || j > 0) {
            uint8_t e = tb[i * len + j];
            uint32_t et;
            if (e == 0) {
                i--;
                j--;
                et = an[j] == bn[i] ? ORC_EDIT_MATCH : ORC_EDIT_MISMATCH;
            } else if (e == 1) {
                i--;
                et = swap ? ORC_EDIT_BGAP : ORC_EDIT_AGAP;
            } else if (e == 2) {
                j--;
                et = swap ? ORC_EDIT_AGAP : ORC_EDIT_BGAP;
            } else {
                i -= 2;
                j -= 2;
                et = ORC_EDIT_TRANSPOSE;
            }
            ev_push(&ev, et);
        }
        ev_reverse(&ev);
        *edits_out = ev.v;
        *n_edits_out = ev.n;
        free(tb);
    }
    free(dp0);
    free(dp1);
    free(dp2);
    free(a_gap);
    free(b_gap);
    return res;
}

uint32_t orc_levenshtein_naive_k_with_opts(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len,
                                           uint32_t k, orc_costs c, orc_edit **edits_out, size_t *n_edits_out) {
    /* src/levenshtein.rs:386-607 */
    int trace_on = edits_out != NULL;
    int swap = a_len > b_len; /* :386 */
    const uint8_t *an = swap ? b : a;
    size_t an_len = swap ? b_len : a_len;
    const uint8_t *bn = swap ? a : b;
    size_t bn_len = swap ? a_len : b_len;
    uint32_t mism = c.mismatch, gap = c.gap, sgap = c.start_gap, tcost = c.transpose;
    int allow_t = c.transpose != 0;

    /* :400-423 upper bound on the cost, in case k is too large */
    uint32_t max_k = min_u32((uint32_t)an_len * mism,
                             (((uint32_t)an_len) << 1) * gap +
                                 (an_len == 0 ? 0 : sgap + (bn_len == an_len ? sgap : 0)));
    max_k = min_u32(k, max_k + (uint32_t)(bn_len - an_len) * gap + (bn_len == an_len ? 0 : sgap));
    /* :426 farthest excursion from the main diagonal */
    size_t unit_k = (size_t)(sat_sub(max_k, sgap) / gap);

    if (bn_len - an_len > unit_k) return ORC_NONE; /* :428-430 */

    size_t len = an_len + 1;
    size_t lo = 0;
    size_t hi = min_sz(unit_k + 1, bn_len + 1);
    size_t prev_lo0, prev_lo1 = 0, prev_hi;
    size_t k_len = min_sz((unit_k << 1) + 1, bn_len + 1);
    uint32_t *dp0 = (uint32_t *)calloc(k_len, 4), *dp1 = (uint32_t *)calloc(k_len, 4),
             *dp2 = (uint32_t *)calloc(k_len, 4);
    uint32_t *a_gap = (uint32_t *)malloc(k_len * 4), *b_gap = (uint32_t *)malloc(k_len * 4);
    for (size_t i = 0; i < k_len; i++) a_gap[i] = b_gap[i] = U32_MAX;
    uint8_t *tb = trace_on ? (uint8_t *)calloc(len * k_len, 1) : NULL;

    for (size_t i = 0; i < hi - lo; i++) { /* :450-456 */
        dp1[i] = (uint32_t)i * gap + (i == 0 ? 0 : sgap);
        if (trace_on) tb[i] = 1;
    }

    for (size_t i = 1; i < len; i++) { /* :458-537 */
        prev_lo0 = prev_lo1;
        prev_lo1 = lo;
        prev_hi = hi;
        hi = min_sz(hi + 1, bn_len + 1);
        if (i > unit_k) lo += 1;

        for (size_t j = 0; j < hi - lo; j++) {
            size_t idx = lo + j;
            uint32_t sub = idx == 0 ? U32_MAX
                                    : dp1[idx - 1 - prev_lo1] + (uint32_t)(an[i - 1] != bn[idx - 1]) * mism;
            a_gap[j] = j == 0 ? U32_MAX : min_u32(dp2[j - 1] + sgap + gap, sat_add(a_gap[j - 1], gap));
            b_gap[j] = idx >= prev_hi
                           ? U32_MAX
                           : min_u32(dp1[idx - prev_lo1] + sgap + gap, sat_add(b_gap[idx - prev_lo1], gap));

            dp2[j] = sub;
            size_t ti = i * k_len + j;
            if (trace_on) tb[ti] = 0;

            if (a_gap[j] < dp2[j]) {
                dp2[j] = a_gap[j];
                if (trace_on) tb[ti] = 1;
            }
            if (b_gap[j] < dp2[j]) {
                dp2[j] = b_gap[j];
                if (trace_on) tb[ti] = 2;
            }
            if (allow_t && i > 1 && idx > 1 && an[i - 1] == bn[idx - 2] && an[i - 2] == bn[idx - 1]) {
                uint32_t tr = dp0[idx - prev_lo0 - 2] + tcost;
                if (tr <= dp2[j]) {
                    dp2[j] = tr;
                    if (trace_on) tb[ti] = 3;
                }
            }
        }
        uint32_t *t = dp0; /* :535-536 */
        dp0 = dp1;
        dp1 = dp2;
        dp2 = t;
    }

    uint32_t res = dp1[hi - lo - 1];
    int none = res > max_k; /* :539-541 */

    if (trace_on && !none) { /* :547-606 */
        edit_vec ev = {NULL, 0, 0};
        size_t i = an_len, j = bn_len;
        while (i > 0 || j > 0) {
            uint8_t e = tb[i * k_len + (j - (i > unit_k ? i - unit_k : 0))];
            uint32_t et;
            if (e == 0) {
                i--;
                j--;
                et = an[i] == bn[j] ? ORC_EDIT_MATCH : ORC_EDIT_MISMATCH;
            } else if (e == 1) {
                j--;
                et = swap ? ORC_EDIT_BGAP : ORC_EDIT_AGAP;
            } else if (e == 2) {
                i--;
                et = swap ? ORC_EDIT_AGAP : ORC_EDIT_BGAP;
            } else {
                i -= 2;
                j -= 2;
                et = ORC_EDIT_TRANSPOSE;
            }
            ev_push(&ev, et);
        }
        ev_reverse(&ev);
        *edits_out = ev.v;
        *n_edits_out = ev.n;
    } else if (trace_on) {
        *edits_out = NULL;
        *n_edits_out = 0;
    }
    free(tb);
    free(dp0);
    free(dp1);
    free(dp2);
    free(a_gap);
    free(b_gap);
    return none ? ORC_NONE : res;
}

uint32_t orc_levenshtein_exp_with_opts(const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len, orc_costs c) {
    /* src/levenshtein.rs:1486-1493: k = 30; loop { k-bounded; k *= 2 }.  The SIMD entry returns Some(0) for two
     * empty strings before anything else (:721-727); the scalar routine gives the same answer. */
    uint32_t k = 30;
    for (;;) {
        uint32_t r = orc_levenshtein_naive_k_with_opts(a, a_len, b, b_len, k, c, NULL, NULL);
        if (r != ORC_NONE) return r;
        k *= 2;
    }
}

uint32_t orc_search_default_k(size_t needle_len) { /* src/levenshtein.rs:1556 */
    return (((uint32_t)needle_len) >> 1) + (((uint32_t)needle_len) & 1);
}

typedef struct {
    orc_match *v;
    size_t n, cap;
} match_vec;
static void mv_push(match_vec *m, uint64_t start, uint64_t end, uint32_t k) {
    if (m->n == m->cap) {
        m->cap = m->cap ? m->cap * 2 : 8;
        m->v = (orc_match *)realloc(m->v, m->cap * sizeof(orc_match));
    }
    m->v[m->n].start = start;
    m->v[m->n].end = end;
    m->v[m->n].k = k;
    m->v[m->n]._pad = 0;
    m->n++;
}

int64_t orc_levenshtein_search_naive_with_opts(const uint8_t *needle, size_t needle_len, const uint8_t *haystack,
                                               size_t haystack_len, uint32_t k, int search_type, orc_costs c,
                                               int anchored, orc_match **out) {
    /* src/levenshtein.rs:1597-1838.  The reference is a lazy iterator; here the whole stream is materialised
     * (emission order and the running Best threshold are the same). */
    match_vec mv = {NULL, 0, 0};
    int best = search_type == 1;
    *out = NULL;

    if (needle_len == 0) { /* :1600-1644 */
        if (!anchored) return 0;
        mv_push(&mv, 0, 0, 0);
        if (!best) {
            /* the iterator stops at the first position whose cost exceeds k (from_fn returns None there) */
            uint32_t cost = c.start_gap;
            for (size_t i = 0; i < haystack_len; i++) {
                cost += c.gap;
                if (cost <= k)
                    mv_push(&mv, 0, i + 1, cost);
                else
                    break;
            }
        }
        *out = mv.v;
        return (int64_t)mv.n;
    }

    if (!orc_costs_valid_search(c)) return -1; /* :1647 */

    size_t len = needle_len + 1;
    size_t iter_len; /* :1650-1661 */
    if (anchored) {
        size_t extra = (size_t)sat_sub(k, c.start_gap) / (size_t)c.gap;
        size_t lim = needle_len + extra;
        if (lim < needle_len) lim = (size_t)-1; /* saturating_add */
        iter_len = min_sz(haystack_len, lim);
    } else {
        iter_len = haystack_len;
    }

    uint32_t *dp0 = (uint32_t *)calloc(len, 4), *dp1 = (uint32_t *)calloc(len, 4), *dp2 = (uint32_t *)calloc(len, 4);
    uint32_t *ngap = (uint32_t *)malloc(len * 4), *hgap = (uint32_t *)malloc(len * 4);
    for (size_t j = 0; j < len; j++) ngap[j] = hgap[j] = U32_MAX;
    size_t *len0 = (size_t *)calloc(len, sizeof(size_t)), *len1 = (size_t *)calloc(len, sizeof(size_t)),
           *len2 = (size_t *)calloc(len, sizeof(size_t));
    size_t *ngap_len = (size_t *)calloc(len, sizeof(size_t)), *hgap_len = (size_t *)calloc(len, sizeof(size_t));
    uint32_t curr_k = k;
    uint32_t mism = c.mismatch, gap = c.gap, sgap = c.start_gap, tcost = c.transpose;
    int allow_t = c.transpose != 0;

    /* :1686-1707 first call: row 0 */
    for (size_t j = 0; j < len; j++) dp1[j] = (uint32_t)j * gap + (j == 0 ? 0 : sgap);
    if (dp1[len - 1] <= curr_k) {
        if (best) curr_k = dp1[len - 1];
        mv_push(&mv, 0, 0, dp1[len - 1]);
    }

    for (size_t i = 0; i < iter_len;) { /* :1709-1807 */
        ngap[0] = anchored ? ((uint32_t)i + 1) * gap + sgap : 0;
        dp2[0] = anchored ? ((uint32_t)i + 1) * gap + sgap : 0;
        ngap_len[0] = 0;
        len2[0] = 0;

        for (size_t j = 1; j < len; j++) {
            uint32_t sub = dp1[j - 1] + (uint32_t)(needle[j - 1] != haystack[i]) * mism;

            uint32_t new_gap = dp1[j] + sgap + gap; /* :1726-1737 */
            uint32_t cont_gap = sat_add(ngap[j], gap);
            if (new_gap < cont_gap) {
                ngap[j] = new_gap;
                ngap_len[j] = len1[j] + 1;
            } else if (new_gap > cont_gap) {
                ngap[j] = cont_gap;
                ngap_len[j] += 1;
            } else {
                ngap[j] = cont_gap;
                ngap_len[j] = max_sz(len1[j], ngap_len[j]) + 1;
            }

            new_gap = dp2[j - 1] + sgap + gap; /* :1739-1750 */
            cont_gap = sat_add(hgap[j - 1], gap);
            if (new_gap < cont_gap) {
                hgap[j] = new_gap;
                hgap_len[j] = len2[j - 1];
            } else if (new_gap > cont_gap) {
                hgap[j] = cont_gap;
                hgap_len[j] = hgap_len[j - 1];
            } else {
                hgap[j] = cont_gap;
                hgap_len[j] = max_sz(len2[j - 1], hgap_len[j - 1]);
            }

            dp2[j] = ngap[j]; /* :1752-1753 */
            len2[j] = ngap_len[j];

            /* :1755-1760 -- compares length2[j-1] (not haystack_gap_length[j]) exactly as the reference does */
            if (hgap[j] < dp2[j] || (hgap[j] == dp2[j] && len2[j - 1] > len2[j])) {
                dp2[j] = hgap[j];
                len2[j] = hgap_len[j];
            }
            if (sub < dp2[j] || (sub == dp2[j] && (len1[j - 1] + 1) > len2[j])) { /* :1762-1765 */
                dp2[j] = sub;
                len2[j] = len1[j - 1] + 1;
            }
            if (allow_t && i > 0 && j > 1 && needle[j - 1] == haystack[i - 1] && needle[j - 2] == haystack[i]) {
                uint32_t tr = dp0[j - 2] + tcost; /* :1767-1779 */
                if (tr <= dp2[j]) {
                    dp2[j] = tr;
                    len2[j] = len0[j - 2] + 2;
                }
            }
        }

        uint32_t final_res = dp2[len - 1];
        size_t final_length = len2[len - 1];

        uint32_t *t = dp0; /* :1785-1788 */
        dp0 = dp1;
        dp1 = dp2;
        dp2 = t;
        size_t *tl = len0;
        len0 = len1;
        len1 = len2;
        len2 = tl;

        i += 1;

        if (final_res <= curr_k) { /* :1792-1806 */
            if (best) curr_k = final_res;
            mv_push(&mv, i - final_length, i, final_res);
        }
    }

    if (best && mv.n > 0) { /* :1812-1835 */
        size_t w = 0;
        for (size_t r = 0; r < mv.n; r++) {
            if (w == 0) {
                mv.v[w++] = mv.v[r];
            } else if (mv.v[r].start <= mv.v[w - 1].start) {
                mv.v[w - 1] = mv.v[r]; /* replace previous if fully overlapping */
            } else {
                mv.v[w++] = mv.v[r];
            }
        }
        size_t f = 0;
        for (size_t r = 0; r < w; r++)
            if (mv.v[r].k == curr_k) mv.v[f++] = mv.v[r];
        mv.n = f;
    }

    free(dp0);
    free(dp1);
    free(dp2);
    free(ngap);
    free(hgap);
    free(len0);
    free(len1);
    free(len2);
    free(ngap_len);
    free(hgap_len);
    *out = mv.v;
    return (int64_t)mv.n;
}

int64_t orc_hamming_search_naive_with_opts(const uint8_t *needle, size_t needle_len, const uint8_t *haystack,
                                           size_t haystack_len, uint32_t k, int search_type, orc_match **out) {
    /* src/hamming.rs:96-146 */
    match_vec mv = {NULL, 0, 0};
    *out = NULL;
    if (needle_len > haystack_len) return 0; /* :100-102 */
    size_t len = haystack_len + 1 - needle_len;
    uint32_t curr_k = k;
    int best = search_type == 1;
    for (size_t i = 0; i < len; i++) { /* :108-131 */
        uint32_t final_res = 0;
        int skip = 0;
        for (size_t j = 0; j < needle_len; j++) {
            final_res += (needle[j] != haystack[i + j]);
            if (final_res > curr_k) { /* early stop */
                skip = 1;
                break;
            }
        }
        if (skip) continue;
        if (best) curr_k = final_res;
        mv_push(&mv, i, i + needle_len, final_res);
    }
    if (best) { /* :136-143 */
        size_t f = 0;
        for (size_t r = 0; r < mv.n; r++)
            if (mv.v[r].k == curr_k) mv.v[f++] = mv.v[r];
        mv.n = f;
    }
    *out = mv.v;
    return (int64_t)mv.n;
}

int64_t orc_hamming_search_with_opts(const uint8_t *needle, size_t needle_len, const uint8_t *haystack,
                                     size_t haystack_len, uint32_t k, int search_type, orc_match **out) {
    /* src/hamming.rs:454-475 */
    *out = NULL;
    if (needle_len > haystack_len) return 0;
    if (needle_len == 0) return 0;
    for (size_t i = 0; i < haystack_len; i++) /* check_no_null_bytes, src/lib.rs:237-243 */
        if (haystack[i] == 0) return -2;
    return orc_hamming_search_naive_with_opts(needle, needle_len, haystack, haystack_len, k, search_type, out);
}

/* ------------------------------------------------------------------------------------------------ */
/* batch drivers for the CPU baseline: a pthread parallel-for with dynamic chunking (libgomp is not   */
/* in this image).  The reference has no batch API or threads; this is "a user's loop over pairs".   */

typedef void (*range_fn)(void *ctx, size_t lo, size_t hi);
typedef struct {
    range_fn fn;
    void *ctx;
    size_t n, chunk;
    size_t next; /* atomic */
} pf_state;

static void *pf_worker(void *p) {
    pf_state *s = (pf_state *)p;
    for (;;) {
        size_t lo = __atomic_fetch_add(&s->next, s->chunk, __ATOMIC_RELAXED);
        if (lo >= s->n) break;
        size_t hi = lo + s->chunk < s->n ? lo + s->chunk : s->n;
        s->fn(s->ctx, lo, hi);
    }
    return NULL;
}

static void parallel_for(size_t n, size_t chunk, int n_threads, range_fn fn, void *ctx) {
    pf_state s = {fn, ctx, n, chunk ? chunk : 1, 0};
    if (n_threads <= 1 || n <= chunk) {
        pf_worker(&s);
        return;
    }
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < n_threads - 1; t++)
        if (pthread_create(&th[started], NULL, pf_worker, &s) == 0) started++;
    pf_worker(&s);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    const uint8_t *a, *b;
    const uint64_t *a_off, *b_off;
    uint32_t k;
    orc_costs c;
    uint32_t *out;
    /* search */
    const uint8_t *needle;
    size_t needle_len;
    int search_type, anchored, simd;
    orc_match **per;
    int64_t *cnt;
} batch_ctx;

static void hamming_range(void *p, size_t lo, size_t hi) {
    batch_ctx *x = (batch_ctx *)p;
    for (size_t i = lo; i < hi; i++) {
        int64_t r = orc_hamming_naive(x->a + x->a_off[i], x->a_off[i + 1] - x->a_off[i], x->b + x->b_off[i],
                                      x->b_off[i + 1] - x->b_off[i]);
        x->out[i] = r < 0 ? ORC_NONE : (uint32_t)r;
    }
}
static void lev_k_range(void *p, size_t lo, size_t hi) {
    batch_ctx *x = (batch_ctx *)p;
    for (size_t i = lo; i < hi; i++)
        x->out[i] = orc_levenshtein_naive_k_with_opts(x->a + x->a_off[i], x->a_off[i + 1] - x->a_off[i],
                                                      x->b + x->b_off[i], x->b_off[i + 1] - x->b_off[i], x->k, x->c,
                                                      NULL, NULL);
}
static void lev_exp_range(void *p, size_t lo, size_t hi) {
    batch_ctx *x = (batch_ctx *)p;
    for (size_t i = lo; i < hi; i++)
        x->out[i] = orc_levenshtein_exp_with_opts(x->a + x->a_off[i], x->a_off[i + 1] - x->a_off[i],
                                                  x->b + x->b_off[i], x->b_off[i + 1] - x->b_off[i], x->c);
}
static void hsearch_range(void *p, size_t lo, size_t hi) {
    batch_ctx *x = (batch_ctx *)p;
    for (size_t i = lo; i < hi; i++)
        x->cnt[i] = orc_hamming_search_with_opts(x->needle, x->needle_len, x->a + x->a_off[i],
                                                 x->a_off[i + 1] - x->a_off[i], x->k, x->search_type, &x->per[i]);
}
static void search_range(void *p, size_t lo, size_t hi) {
    batch_ctx *x = (batch_ctx *)p;
    for (size_t i = lo; i < hi; i++)
        x->cnt[i] = x->simd ? orc_levenshtein_search_simd_with_opts(x->needle, x->needle_len, x->a + x->a_off[i],
                                                                    x->a_off[i + 1] - x->a_off[i], x->k, x->search_type,
                                                                    x->c, x->anchored, &x->per[i], NULL)
                            : orc_levenshtein_search_naive_with_opts(x->needle, x->needle_len, x->a + x->a_off[i],
                                                                     x->a_off[i + 1] - x->a_off[i], x->k,
                                                                     x->search_type, x->c, x->anchored, &x->per[i]);
}

/* ---- batch drivers over the AVX2 restatement of the reference's SIMD path (ta_ref_avx2.c): CPU baseline only ---- */
static void lev_simd_k_range(void *p, size_t lo, size_t hi) {
    batch_ctx *x = (batch_ctx *)p;
    for (size_t i = lo; i < hi; i++)
        x->out[i] = orc_levenshtein_simd_k_with_opts(x->a + x->a_off[i], x->a_off[i + 1] - x->a_off[i],
                                                     x->b + x->b_off[i], x->b_off[i + 1] - x->b_off[i], x->k, x->c, NULL);
}
static void lev_simd_exp_range(void *p, size_t lo, size_t hi) {
    batch_ctx *x = (batch_ctx *)p;
    for (size_t i = lo; i < hi; i++)
        x->out[i] = orc_levenshtein_simd_exp_with_opts(x->a + x->a_off[i], x->a_off[i + 1] - x->a_off[i],
                                                       x->b + x->b_off[i], x->b_off[i + 1] - x->b_off[i], x->c);
}
static void hamming_simd_range(void *p, size_t lo, size_t hi) {
    batch_ctx *x = (batch_ctx *)p;
    for (size_t i = lo; i < hi; i++) {
        int64_t r = orc_hamming_simd(x->a + x->a_off[i], x->a_off[i + 1] - x->a_off[i], x->b + x->b_off[i],
                                     x->b_off[i + 1] - x->b_off[i]);
        x->out[i] = r < 0 ? ORC_NONE : (uint32_t)r;
    }
}
void orc_levenshtein_simd_k_batch(const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                                  size_t n, uint32_t k, orc_costs c, uint32_t *out, int n_threads) {
    batch_ctx x = {0};
    x.a = a, x.b = b, x.a_off = a_off, x.b_off = b_off, x.out = out, x.k = k, x.c = c;
    parallel_for(n, 256, n_threads, lev_simd_k_range, &x);
}
void orc_levenshtein_simd_exp_batch(const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                                    size_t n, orc_costs c, uint32_t *out, int n_threads) {
    batch_ctx x = {0};
    x.a = a, x.b = b, x.a_off = a_off, x.b_off = b_off, x.out = out, x.c = c;
    parallel_for(n, 64, n_threads, lev_simd_exp_range, &x);
}
void orc_hamming_simd_batch(const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off, size_t n,
                            uint32_t *out, int n_threads) {
    batch_ctx x = {0};
    x.a = a, x.b = b, x.a_off = a_off, x.b_off = b_off, x.out = out;
    parallel_for(n, 1024, n_threads, hamming_simd_range, &x);
}

void orc_hamming_batch(const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off, size_t n,
                       uint32_t *out, int n_threads) {
    batch_ctx x = {0};
    x.a = a, x.b = b, x.a_off = a_off, x.b_off = b_off, x.out = out;
    parallel_for(n, 1024, n_threads, hamming_range, &x);
}

void orc_levenshtein_k_batch(const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                             size_t n, uint32_t k, orc_costs c, uint32_t *out, int n_threads) {
    batch_ctx x = {0};
    x.a = a, x.b = b, x.a_off = a_off, x.b_off = b_off, x.out = out, x.k = k, x.c = c;
    parallel_for(n, 256, n_threads, lev_k_range, &x);
}

void orc_levenshtein_exp_batch(const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                               size_t n, orc_costs c, uint32_t *out, int n_threads) {
    batch_ctx x = {0};
    x.a = a, x.b = b, x.a_off = a_off, x.b_off = b_off, x.out = out, x.c = c;
    parallel_for(n, 64, n_threads, lev_exp_range, &x);
}

static int64_t search_batch_impl(const uint8_t *needle, size_t needle_len, const uint8_t *hay, const uint64_t *hay_off,
                                 size_t n, uint32_t k, int search_type, orc_costs c, int anchored, orc_match **out,
                                 uint64_t *match_off, int n_threads, int simd);
int64_t orc_levenshtein_search_batch(const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                                     const uint64_t *hay_off, size_t n, uint32_t k, int search_type, orc_costs c,
                                     int anchored, orc_match **out, uint64_t *match_off, int n_threads) {
    return search_batch_impl(needle, needle_len, hay, hay_off, n, k, search_type, c, anchored, out, match_off, n_threads, 0);
}
/* the same loop over the AVX2 restatement of the reference's SIMD search (ta_ref_avx2.c): CPU baseline only */
int64_t orc_levenshtein_search_simd_batch(const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                                          const uint64_t *hay_off, size_t n, uint32_t k, int search_type, orc_costs c,
                                          int anchored, orc_match **out, uint64_t *match_off, int n_threads) {
    return search_batch_impl(needle, needle_len, hay, hay_off, n, k, search_type, c, anchored, out, match_off, n_threads, 1);
}
static int64_t search_batch_impl(const uint8_t *needle, size_t needle_len, const uint8_t *hay, const uint64_t *hay_off,
                                 size_t n, uint32_t k, int search_type, orc_costs c, int anchored, orc_match **out,
                                 uint64_t *match_off, int n_threads, int simd) {
    batch_ctx x = {0};
    x.a = hay, x.a_off = hay_off, x.k = k, x.c = c, x.needle = needle, x.needle_len = needle_len;
    x.search_type = search_type, x.anchored = anchored, x.simd = simd;
    x.per = (orc_match **)calloc(n ? n : 1, sizeof(orc_match *));
    x.cnt = (int64_t *)calloc(n ? n : 1, sizeof(int64_t));
    parallel_for(n, 16, n_threads, search_range, &x);
    int bad = 0;
    for (size_t i = 0; i < n; i++)
        if (x.cnt[i] < 0) bad = 1;
    int64_t total = 0;
    if (!bad) {
        match_off[0] = 0;
        for (size_t i = 0; i < n; i++) {
            total += x.cnt[i];
            match_off[i + 1] = (uint64_t)total;
        }
        *out = (orc_match *)malloc((size_t)(total ? total : 1) * sizeof(orc_match));
        for (size_t i = 0; i < n; i++)
            if (x.cnt[i] > 0) memcpy(*out + match_off[i], x.per[i], (size_t)x.cnt[i] * sizeof(orc_match));
    } else {
        *out = NULL;
        total = -1;
    }
    for (size_t i = 0; i < n; i++) free(x.per[i]);
    free(x.per);
    free(x.cnt);
    return total;
}

int64_t orc_hamming_search_batch(const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                                 const uint64_t *hay_off, size_t n, uint32_t k, int search_type, orc_match **out,
                                 uint64_t *match_off, int n_threads) {
    batch_ctx x = {0};
    x.a = hay, x.a_off = hay_off, x.k = k, x.needle = needle, x.needle_len = needle_len;
    x.search_type = search_type;
    x.per = (orc_match **)calloc(n ? n : 1, sizeof(orc_match *));
    x.cnt = (int64_t *)calloc(n ? n : 1, sizeof(int64_t));
    parallel_for(n, 16, n_threads, hsearch_range, &x);
    int64_t bad = 0, total = 0;
    for (size_t i = 0; i < n; i++)
        if (x.cnt[i] < 0) bad = x.cnt[i];
    if (!bad) {
        match_off[0] = 0;
        for (size_t i = 0; i < n; i++) {
            total += x.cnt[i];
            match_off[i + 1] = (uint64_t)total;
        }
        *out = (orc_match *)malloc((size_t)(total ? total : 1) * sizeof(orc_match));
        for (size_t i = 0; i < n; i++)
            if (x.cnt[i] > 0) memcpy(*out + match_off[i], x.per[i], (size_t)x.cnt[i] * sizeof(orc_match));
    } else {
        *out = NULL;
        total = bad;
    }
    for (size_t i = 0; i < n; i++) free(x.per[i]);
    free(x.per);
    free(x.cnt);
    return total;
}

void orc_free(void *p) { free(p); }
