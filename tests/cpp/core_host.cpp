// Host-side differential test of the per-pair bit-parallel cores (triple_accel_b200/csrc/lev_bitpar_core.cuh is
// host/device code) against the scalar oracle.  Test infrastructure: built and run by tests/test_core_host.py.
// usage: core_host <iterations> <seed> [longest string, default 300]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../triple_accel_b200/csrc/lev_bitpar_core.cuh"
extern "C" {
#include "../../oracle/ta_oracle.h"
}

static long bad = 0;

static void report(const char *what, int trans, uint32_t k, int la, int lb, uint32_t want, uint32_t got) {
    if (bad++ < 20) printf("%s MISMATCH trans=%d k=%u la=%d lb=%d want=%u got=%u\n", what, trans, k, la, lb, want, got);
}

template <typename T, size_t N>
static void check_clean(const char *what, T (&t)[N]) {
    for (size_t q = 0; q < N; q++)
        if (t[q]) {
            if (bad++ < 20) printf("%s table not clean at %zu\n", what, q);
            t[q] = 0;
        }
}

// block-table variants
template <bool TRANS, int PLANES, int C>
static uint32_t run_blk(const uint8_t *a, int la, const uint8_t *b, int lb, uint32_t k) {
    static uint32_t tab[256];
    const uint32_t got = bitpar::pair_unit_costs_blk<TRANS, PLANES, C>(a, la, b, lb, k, (uint8_t *)tab, 4u);
    check_clean("blk", tab);
    return got;
}

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 100000;
    std::mt19937_64 rng(argc > 2 ? atoll(argv[2]) : 12345);
    const int long_len = argc > 3 ? atoi(argv[3]) : 300;
    std::vector<uint8_t> arena(1 << 16);  // strings live at base + {0, 1024, 4096, 6144} + small offsets
    uint8_t *base = (uint8_t *)(((uintptr_t)arena.data() + 4096) & ~(uintptr_t)15);
    long tests = 0, duo_tests = 0;
    for (int it = 0; it < iters; it++) {
        const int alphas[5] = {2, 3, 4, 26, 256};
        const int alpha = alphas[rng() % 5];
        const int maxlen = it % 3 == 0 ? long_len : 40;
        int la = (int)(rng() % maxlen), lb;
        const int mode = (int)(rng() % 3);
        const size_t offa = rng() % 64, offb = 1024 + rng() % 64;
        uint8_t *a = base + offa, *b = base + offb;
        for (int i = 0; i < 8192; i++) base[i] = (uint8_t)rng();  // junk around the strings
        for (int i = 0; i < la; i++) a[i] = (uint8_t)(rng() % alpha);
        if (mode == 0) {
            lb = (int)(rng() % maxlen);
            for (int i = 0; i < lb; i++) b[i] = (uint8_t)(rng() % alpha);
        } else {
            std::vector<uint8_t> s(a, a + la);
            const int ne = (int)(rng() % (it % 5 == 0 ? 70 : 12));
            for (int e = 0; e < ne; e++) {
                const int kind = (int)(rng() % 4);
                if (kind == 0 && !s.empty())
                    s[rng() % s.size()] = (uint8_t)(rng() % alpha);
                else if (kind == 1)
                    s.insert(s.begin() + rng() % (s.size() + 1), (uint8_t)(rng() % alpha));
                else if (kind == 2 && !s.empty())
                    s.erase(s.begin() + rng() % s.size());
                else if (kind == 3 && s.size() > 1) {
                    const size_t p = rng() % (s.size() - 1);
                    std::swap(s[p], s[p + 1]);
                }
            }
            lb = (int)s.size();
            memcpy(b, s.data(), lb);
        }
        for (int trans = 0; trans < 2; trans++) {
            const orc_costs c = {1, 1, 0, (uint8_t)trans};
            auto want_for = [&](uint32_t k) { return orc_levenshtein_naive_k_with_opts(a, la, b, lb, k, c, NULL, NULL); };
            {  // 32-row sliding table (1 and 2 planes) and the SWAR kernel
                const uint32_t k = (uint32_t)(rng() % (trans ? 31 : 32));
                const uint32_t want = want_for(k);
                static uint32_t tab[128];
                uint32_t got;
                if (it & 1)
                    got = trans ? bitpar::pair_unit_costs_tab<true, 1, uint32_t>(a, la, b, lb, k, (uint8_t *)tab, 4)
                                : bitpar::pair_unit_costs_tab<false, 1, uint32_t>(a, la, b, lb, k, (uint8_t *)tab, 4);
                else
                    got = trans ? bitpar::pair_unit_costs_tab<true, 2, uint32_t>(a, la, b, lb, k, (uint8_t *)tab, 4)
                                : bitpar::pair_unit_costs_tab<false, 2, uint32_t>(a, la, b, lb, k, (uint8_t *)tab, 4);
                check_clean("tab32", tab);
                if (got != want) report("TAB32", trans, k, la, lb, want, got);
                const uint32_t gs = trans ? bitpar::pair_unit_costs<true>(a, la, b, lb, k)
                                          : bitpar::pair_unit_costs<false>(a, la, b, lb, k);
                if (gs != want) report("SWAR", trans, k, la, lb, want, gs);
            }
            {  // 16-row sliding table
                static uint16_t tab16[128];
                const uint32_t k = (uint32_t)(rng() % (trans ? 15 : 16));
                const uint32_t want = want_for(k);
                const uint32_t got = trans ? bitpar::pair_unit_costs_tab<true, 1, uint16_t>(a, la, b, lb, k, (uint8_t *)tab16, 2)
                                           : bitpar::pair_unit_costs_tab<false, 1, uint16_t>(a, la, b, lb, k, (uint8_t *)tab16, 2);
                check_clean("tab16", tab16);
                if (got != want) report("TAB16", trans, k, la, lb, want, got);
            }
            {  // 64-row sliding table
                static uint64_t tab64[64];
                const uint32_t k = (uint32_t)(rng() % (trans ? 63 : 64));
                const uint32_t want = want_for(k);
                const uint32_t got = trans ? bitpar::pair_unit_costs_tab<true, 2, uint64_t>(a, la, b, lb, k, (uint8_t *)tab64, 8)
                                           : bitpar::pair_unit_costs_tab<false, 2, uint64_t>(a, la, b, lb, k, (uint8_t *)tab64, 8);
                check_clean("tab64", tab64);
                if (got != want) report("TAB64", trans, k, la, lb, want, got);
            }
            {  // block table, 16-position blocks: bands of <= 17 diagonals
                const uint32_t k = (uint32_t)(rng() % (trans ? 16 : 17));
                const uint32_t want = want_for(k);
                uint32_t got;
                if ((it >> 1) & 1)
                    got = trans ? run_blk<true, 1, 16>(a, la, b, lb, k) : run_blk<false, 1, 16>(a, la, b, lb, k);
                else
                    got = trans ? run_blk<true, 0, 16>(a, la, b, lb, k) : run_blk<false, 0, 16>(a, la, b, lb, k);
                if (got != want) report("BLK16", trans, k, la, lb, want, got);
            }
            {  // block table, 8-position blocks: bands of <= 25 diagonals
                const uint32_t k = (uint32_t)(rng() % (trans ? 24 : 25));
                const uint32_t want = want_for(k);
                uint32_t got;
                if ((it >> 1) & 1)
                    got = trans ? run_blk<true, 1, 8>(a, la, b, lb, k) : run_blk<false, 1, 8>(a, la, b, lb, k);
                else
                    got = trans ? run_blk<true, 0, 8>(a, la, b, lb, k) : run_blk<false, 0, 8>(a, la, b, lb, k);
                if (got != want) report("BLK8", trans, k, la, lb, want, got);
            }
            tests++;
        }
        {  // two pairs per thread (bands <= 9 diagonals): this iteration's pair + a partner with the same superstep count
            // partner = a mutation of this pair with similar lengths
            std::vector<uint8_t> qa(a, a + la), qb(b, b + lb);
            for (int rep = 0; rep < 2; rep++) {
                std::vector<uint8_t> &v = rep ? qb : qa;
                const int ne = (int)(rng() % 4);
                for (int e = 0; e < ne; e++) {
                    const int kind = (int)(rng() % 3);
                    if (kind == 0 && !v.empty()) v[rng() % v.size()] = (uint8_t)(rng() % alpha);
                    else if (kind == 1) v.insert(v.begin() + rng() % (v.size() + 1), (uint8_t)(rng() % alpha));
                    else if (!v.empty()) v.erase(v.begin() + rng() % v.size());
                }
            }
            uint8_t *a2 = base + 4096 + rng() % 64, *b2 = base + 6144 + rng() % 64;
            memcpy(a2, qa.data(), qa.size());
            memcpy(b2, qb.data(), qb.size());
            const uint32_t k = (uint32_t)(rng() % 9);
            const orc_costs c = {1, 1, 0, 0};
            const uint32_t want1 = orc_levenshtein_naive_k_with_opts(a, la, b, lb, k, c, NULL, NULL);
            const uint32_t want2 = orc_levenshtein_naive_k_with_opts(a2, qa.size(), b2, qb.size(), k, c, NULL, NULL);
            const uint8_t *x1 = a, *y1 = b, *x2 = a2, *y2 = b2;
            uint64_t lx1 = la, ly1 = lb, lx2 = qa.size(), ly2 = qb.size();
            uint32_t mk1, mk2, o1 = 0, o2 = 0;
            const bool dp1 = bitpar::unit_costs_prepare(x1, lx1, y1, ly1, k, mk1, &o1);
            const bool dp2 = bitpar::unit_costs_prepare(x2, lx2, y2, ly2, k, mk2, &o2);
            if (dp1 && dp2 && (ly1 >> 4) == (ly2 >> 4)) {
                static uint32_t tab[128];
                uint32_t d1, d2;
                bitpar::distance_duo(x1, (int)lx1, y1, (int)ly1, mk1, x2, (int)lx2, y2, (int)ly2, mk2, (uint8_t *)tab, 4u, d1, d2);
                check_clean("duo", tab);
                d1 = d1 <= mk1 ? d1 : 0xFFFFFFFFu;
                d2 = d2 <= mk2 ? d2 : 0xFFFFFFFFu;
                if (d1 != want1) report("DUO-A", 0, k, la, lb, want1, d1);
                if (d2 != want2) report("DUO-B", 0, k, (int)qa.size(), (int)qb.size(), want2, d2);
                duo_tests++;
            } else {
                if (!dp1 && o1 != want1) report("PREP-A", 0, k, la, lb, want1, o1);
                if (!dp2 && o2 != want2) report("PREP-B", 0, k, (int)qa.size(), (int)qb.size(), want2, o2);
            }
        }
    }
    printf("tests %ld (duo %ld) bad %ld\n", tests, duo_tests, bad);
    return bad != 0;
}
