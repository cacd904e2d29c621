// Host-only test of the search path's host phase (triple_accel_b200/csrc/search_emit.hpp: hit ordering + the
// reference's emission rules) against the scalar oracle.  Test infrastructure: built and run by
// tests/test_search_emit_host.py.  The device phase is replaced by the oracle's SearchType::All output -- every end
// position with cost <= k is exactly what the exact kernels report -- shuffled, because the kernels append hits in no
// particular order.
// usage: search_emit_host <haystacks> <seed>      differential test
//        search_emit_host bench <haystacks> <hits> time emit_matches on a cfg-4-like hit list
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../triple_accel_b200/csrc/search_emit.hpp"
extern "C" {
#include "../../oracle/ta_oracle.h"
}

static int run_bench(size_t n, size_t nh) {
    std::mt19937_64 rng(7);
    std::vector<Hit> base;
    for (size_t i = 0; i < nh; i++) {
        const uint32_t h = (uint32_t)(rng() % n) & ~3u;  // ~4 hits per matching haystack
        base.push_back(Hit{h, (uint32_t)(100 + rng() % 3000), 32, (uint32_t)(rng() % 4)});
    }
    std::vector<uint64_t> moff(n + 1);
    double best_us = 1e30;
    size_t matches = 0;
    for (int rep = 0; rep < 20; rep++) {
        std::vector<Hit> hits = base;
        std::vector<ta_match> result;
        moff[0] = 0;
        const auto t0 = std::chrono::steady_clock::now();
        emit_matches(n, 32, 3, true, ta_costs{1, 1, 0, 0}, hits, moff.data(), result);
        const auto t1 = std::chrono::steady_clock::now();
        best_us = std::min(best_us, (double)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count() / 1e3);
        matches = result.size();
    }
    printf("emit_matches: %zu haystacks, %zu hits -> %zu matches, best of 20: %.1f us\n", n, nh, matches, best_us);
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 4 && !strcmp(argv[1], "bench")) return run_bench((size_t)atol(argv[2]), (size_t)atol(argv[3]));
    const size_t n = argc > 1 ? (size_t)atol(argv[1]) : 300;
    std::mt19937_64 rng(argc > 2 ? (uint64_t)atol(argv[2]) : 1);
    const ta_costs cost_set[] = {{1, 1, 0, 0}, {1, 1, 0, 1}, {2, 1, 3, 0}, {1, 2, 0, 1}, {3, 2, 1, 2}, {1, 1, 2, 0}};
    size_t bad = 0, cases = 0, total_matches = 0;
    for (int round = 0; round < 40; round++) {
        const ta_costs costs = cost_set[round % 6];
        const orc_costs oc = {costs.mismatch, costs.gap, costs.start_gap, costs.transpose};
        const uint32_t alpha = 2 + (uint32_t)(rng() % 3);
        const size_t nlen = 1 + rng() % 20;
        std::vector<uint8_t> needle(nlen);
        for (auto &c : needle) c = (uint8_t)(rng() % alpha);
        // k from 0 up to beyond the row-0 threshold (needle_len * gap + start_gap), where every haystack reports (0, 0)
        const uint32_t row0 = (uint32_t)nlen * costs.gap + costs.start_gap;
        const uint32_t k = (uint32_t)(rng() % (row0 + 3));
        std::vector<std::vector<uint8_t>> hays(n);
        for (auto &h : hays) {
            h.resize(rng() % 5 == 0 ? 0 : rng() % 120);
            for (auto &c : h) c = (uint8_t)(rng() % alpha);
            if (h.size() > nlen && rng() % 2) memcpy(h.data() + rng() % (h.size() - nlen), needle.data(), nlen);
        }
        std::vector<Hit> hits;
        for (size_t i = 0; i < n; i++) {
            orc_match *m = nullptr;
            const int64_t cnt = orc_levenshtein_search_naive_with_opts(needle.data(), nlen, hays[i].data(), hays[i].size(), k,
                                                                       0 /* All */, oc, 0, &m);
            for (int64_t j = 0; j < cnt; j++)
                if (m[j].end > 0) hits.push_back(Hit{(uint32_t)i, (uint32_t)m[j].end, (uint32_t)(m[j].end - m[j].start), m[j].k});
            if (m) orc_free(m);
        }
        std::shuffle(hits.begin(), hits.end(), rng);
        for (int best = 0; best < 2; best++) {
            std::vector<Hit> h2 = hits;
            std::vector<uint64_t> moff(n + 1, 0xDEADBEEF);
            moff[0] = 0;
            std::vector<ta_match> result;
            emit_matches(n, nlen, k, best != 0, costs, h2, moff.data(), result);
            for (size_t i = 0; i < n; i++) {
                orc_match *m = nullptr;
                const int64_t cnt = orc_levenshtein_search_naive_with_opts(needle.data(), nlen, hays[i].data(), hays[i].size(),
                                                                           k, best, oc, 0, &m);
                bool ok = moff[i + 1] >= moff[i] && (int64_t)(moff[i + 1] - moff[i]) == cnt && moff[i + 1] <= result.size();
                for (int64_t j = 0; ok && j < cnt; j++) {
                    const ta_match &r = result[moff[i] + j];
                    ok = r.start == m[j].start && r.end == m[j].end && r.k == m[j].k;
                }
                if (m) orc_free(m);
                cases++;
                total_matches += (size_t)cnt;
                if (!ok && bad++ < 5) printf("MISMATCH round %d best %d haystack %zu (k %u, needle %zu)\n", round, best, i, k, nlen);
            }
        }
    }
    printf("search_emit_host: cases %zu matches %zu bad %zu\n", cases, total_matches, bad);
    return bad ? 1 : 0;
}
