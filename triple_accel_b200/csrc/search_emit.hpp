// search_emit.hpp -- host phase of the search path: plain C++ (no CUDA), included by search.cu and compiled on its own
// by tests/cpp/search_emit_host.cpp, which pins the emission rules to the oracle without a GPU.
//
// The device phase returns every end position whose cost is <= k as a Hit (unordered); this file orders them and applies
// the reference's emission rules: the row-0 match (src/levenshtein.rs:1686-1707), the running Best threshold
// (:1792-1806) and the Best post-pass (:1812-1835).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "../../include/triple_accel_b200.h"

struct Hit {  // one reported end position (string lengths are < 2^22, TA_MAX_STRING_LEN)
    uint32_t hay, end, len, cost;
};

// orders the hits by (haystack, end): ONE counting pass over buckets that are monotone in the haystack index
// (bucket = floor(hay * B / n_hay), B = the power of two >= the number of hits, so a bucket holds the hits of about one
// haystack), then each bucket -- a handful of hits as a rule -- by insertion.  History: a comparison sort of a few
// thousand 16-byte records cost more than the exact kernel that produced them (0.10 ms); an LSD radix sort of the
// 64-bit key took three passes over 2048 counters; two 11-bit passes on the haystack index alone 37 us; this is ~12 us.
static inline void sort_hits(std::vector<Hit> &hits, size_t n_hay) {
    const size_t n = hits.size();
    if (n < 2) return;
    auto less = [](const Hit &x, const Hit &y) { return x.hay != y.hay ? x.hay < y.hay : x.end < y.end; };
    if (n < 64 || n_hay == 0) {
        std::sort(hits.begin(), hits.end(), less);
        return;
    }
    size_t B = 256;
    while (B < n && B < (1u << 16)) B <<= 1;
    const uint64_t mul = ((uint64_t)B << 32) / n_hay;  // hay < n_hay <= 2^32: hay * mul < B << 32, monotone in hay
    static thread_local std::vector<uint32_t> count;
    static thread_local std::vector<Hit> tmp;
    count.assign(B + 1, 0);
    tmp.resize(n);
    const Hit *src = hits.data();
    for (size_t i = 0; i < n; i++) count[(size_t)(((uint64_t)src[i].hay * mul) >> 32) + 1]++;
    for (size_t b = 0; b < B; b++) count[b + 1] += count[b];  // count[b] = first slot of bucket b
    Hit *dst = tmp.data();
    for (size_t i = 0; i < n; i++) dst[count[(size_t)(((uint64_t)src[i].hay * mul) >> 32)]++] = src[i];
    // count[b] is now the END of bucket b; buckets are in haystack order, so insertion over the whole array only ever
    // moves a record inside its own bucket
    for (size_t x = 1; x < n; x++) {
        if (!less(dst[x], dst[x - 1])) continue;
        const Hit h = dst[x];
        size_t y = x;
        for (; y > 0 && less(h, dst[y - 1]); y--) dst[y] = dst[y - 1];
        dst[y] = h;
    }
    hits.swap(tmp);
}

// moff[a .. b) = v.  The offsets of the haystacks without a match are ~1000 runs of ~100 equal values per call; std::fill
// pays its alignment prologue / remainder epilogue on every run and was 3x slower than one flat fill of the same bytes.
// Here: groups of eight stores, and the last group overlaps the one before it instead of a scalar tail.
static inline void fill_run(uint64_t *a, uint64_t *b, uint64_t v) {
    if (b - a < 8) {
        for (; a < b; a++) *a = v;
        return;
    }
    for (; a + 8 <= b; a += 8) {
        a[0] = v, a[1] = v, a[2] = v, a[3] = v;
        a[4] = v, a[5] = v, a[6] = v, a[7] = v;
    }
    a = b - 8;
    a[0] = v, a[1] = v, a[2] = v, a[3] = v;
    a[4] = v, a[5] = v, a[6] = v, a[7] = v;
}

// Host phase: order the hits by (haystack, end) and apply the reference's emission rules -- the row-0 match
// (src/levenshtein.rs:1686-1707), the running Best threshold (:1792-1806) and the Best post-pass (:1812-1835).
// moff[0] is set by the caller; moff[1 .. n] and `result` are filled here.
static inline void emit_matches(size_t n, size_t needle_len, uint32_t k, bool best, ta_costs costs, std::vector<Hit> &hits,
                                uint64_t *moff, std::vector<ta_match> &result) {
    static const bool trace = getenv("TA_TRACE_SEARCH") != nullptr;
    const auto t_sort0 = std::chrono::steady_clock::now();
    sort_hits(hits, n);
    if (trace)
        fprintf(stderr, "[ta search] sort %.1f us\n",
                (double)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_sort0).count() / 1e3);
    const uint32_t row0 = (uint32_t)needle_len * costs.gap + costs.start_gap;
    const size_t nh = hits.size();
    const Hit *hv = hits.data();
    size_t hp = 0;
    result.reserve(nh + (row0 <= k ? n : 0));
    for (size_t i = 0; i < n;) {
        if (row0 > k) {  // a run of haystacks with nothing to report: filled at memset speed
            const size_t stop = hp < nh ? (size_t)hv[hp].hay : n;
            if (stop > i) {
                fill_run(moff + i + 1, moff + stop + 1, (uint64_t)result.size());
                i = stop;
                if (i >= n) break;
            }
        }
        const size_t base = result.size();  // this haystack's matches go straight into `result`
        uint32_t curr_k = k;
        if (row0 <= curr_k) {
            if (best) curr_k = row0;
            result.push_back(ta_match{0, 0, row0, 0});
        }
        for (; hp < nh && hv[hp].hay == i; hp++) {
            const Hit &h = hv[hp];
            if (h.cost <= curr_k) {
                if (best) curr_k = h.cost;
                result.push_back(ta_match{h.end - h.len, h.end, h.cost, 0});
            }
        }
        if (best && result.size() > base) {
            ta_match *cur = result.data() + base;
            const size_t cnt = result.size() - base;
            size_t wpos = 0;
            for (size_t r = 0; r < cnt; r++) {
                if (wpos == 0)
                    cur[wpos++] = cur[r];
                else if (cur[r].start <= cur[wpos - 1].start)
                    cur[wpos - 1] = cur[r];  // replace previous if fully overlapping
                else
                    cur[wpos++] = cur[r];
            }
            size_t f = 0;
            for (size_t r = 0; r < wpos; r++)
                if (cur[r].k == curr_k) cur[f++] = cur[r];
            result.resize(base + f);
        }
        moff[++i] = result.size();
    }
}
