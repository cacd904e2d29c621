// Host-side differential test of the thread-per-pair u16 routine (triple_accel_b200/csrc/lev_diag16_core.cuh is
// host/device code; the DPX / PRMT / funnel-shift instructions are emulated on the host) against the scalar oracle.
// Test infrastructure: built and run by tests/test_diag16_host.py.   usage: diag16_host <iterations> <seed>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../triple_accel_b200/csrc/lev_diag16_core.cuh"
extern "C" {
#include "../../oracle/ta_oracle.h"
}

template <int NR>
static uint32_t run(bool affine, bool trans, const uint8_t *a, int la, const uint8_t *b, int lb, uint32_t k, orc_costs c) {
    if (affine && trans) return diag16::pair<NR, true, true>(a, la, b, lb, k, c.mismatch, c.gap, c.start_gap, c.transpose);
    if (affine) return diag16::pair<NR, true, false>(a, la, b, lb, k, c.mismatch, c.gap, c.start_gap, c.transpose);
    if (trans) return diag16::pair<NR, false, true>(a, la, b, lb, k, c.mismatch, c.gap, c.start_gap, c.transpose);
    return diag16::pair<NR, false, false>(a, la, b, lb, k, c.mismatch, c.gap, c.start_gap, c.transpose);
}

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 100000;
    std::mt19937_64 rng(argc > 2 ? atoll(argv[2]) : 12345);
    std::vector<uint8_t> arena(1 << 16);
    uint8_t *base = (uint8_t *)(((uintptr_t)arena.data() + 4096) & ~(uintptr_t)15);
    const orc_costs models[10] = {{1, 1, 0, 0}, {1, 1, 0, 1}, {1, 1, 2, 0}, {2, 1, 2, 0}, {2, 3, 0, 0},
                                  {3, 1, 0, 0}, {2, 2, 1, 3}, {3, 2, 0, 2}, {1, 1, 1, 1}, {5, 4, 3, 0}};
    long tests = 0, bad = 0, skipped = 0, within = 0;
    for (int it = 0; it < iters; it++) {
        const int alphas[5] = {2, 3, 4, 26, 256};
        const int alpha = alphas[rng() % 5];
        const int maxlen = it % 4 == 0 ? 200 : 40;
        int la = (int)(rng() % maxlen), lb;
        const size_t offa = rng() % 64, offb = 1024 + rng() % 64;
        uint8_t *a = base + offa, *b = base + offb;
        for (int i = 0; i < 4096; i++) base[i] = (uint8_t)rng();  // junk around the strings
        for (int i = 0; i < la; i++) a[i] = (uint8_t)(rng() % alpha);
        if (rng() % 3 == 0) {
            lb = (int)(rng() % maxlen);
            for (int i = 0; i < lb; i++) b[i] = (uint8_t)(rng() % alpha);
        } else {
            std::vector<uint8_t> s(a, a + la);
            const int e = (int)(rng() % 8);
            for (int q = 0; q < e; q++) {
                const int kind = (int)(rng() % 4);
                if (kind == 0 && !s.empty()) s[rng() % s.size()] = (uint8_t)(rng() % alpha);
                else if (kind == 1) s.insert(s.begin() + rng() % (s.size() + 1), (uint8_t)(rng() % alpha));
                else if (kind == 2 && !s.empty()) s.erase(s.begin() + rng() % s.size());
                else if (kind == 3 && s.size() > 1) { const size_t p = rng() % (s.size() - 1); std::swap(s[p], s[p + 1]); }
            }
            lb = (int)s.size();
            memcpy(b, s.data(), s.size());
        }
        const orc_costs c = models[rng() % 10];
        const uint32_t ks[8] = {0, 1, 3, 6, 10, 16, 25, 40};
        const uint32_t k = ks[rng() % 8];
        // the band this pair needs (the dispatcher bounds it for the whole batch)
        const int m = la < lb ? la : lb, n = la < lb ? lb : la;
        const BandInfo bi = band_info(m, n, k, c.mismatch, c.gap, c.start_gap, c.transpose != 0);
        if (!bi.none && m > 0 && bi.W > 32) { skipped++; continue; }
        const int W = (bi.none || m == 0) ? 1 : bi.W;
        const uint32_t want = orc_levenshtein_naive_k_with_opts(a, la, b, lb, k, c, nullptr, nullptr);
        const bool affine = c.start_gap != 0, trans = c.transpose != 0;
        // every register count that holds the band, affine code path also for start_gap == 0
        const int nrs[5] = {2, 3, 4, 6, 8};
        for (int q = 0; q < 5; q++) {
            if (4 * nrs[q] < W) continue;
            for (int fa = affine ? 1 : 0; fa < 2; fa++) {
                uint32_t got;
                switch (nrs[q]) {
                    case 2: got = run<2>(fa, trans, a, la, b, lb, k, c); break;
                    case 3: got = run<3>(fa, trans, a, la, b, lb, k, c); break;
                    case 4: got = run<4>(fa, trans, a, la, b, lb, k, c); break;
                    case 6: got = run<6>(fa, trans, a, la, b, lb, k, c); break;
                    default: got = run<8>(fa, trans, a, la, b, lb, k, c); break;
                }
                tests++;
                within += want != 0xFFFFFFFFu;
                if (got != want && bad++ < 20)
                    printf("MISMATCH NR=%d affine_path=%d costs=(%d,%d,%d,%d) k=%u la=%d lb=%d offa=%zu offb=%zu want=%u got=%u\n",
                           nrs[q], fa, c.mismatch, c.gap, c.start_gap, c.transpose, k, la, lb, offa, offb, want, got);
            }
        }
    }
    printf("tests %ld (within k: %ld) skipped %ld bad %ld\n", tests, within, skipped, bad);
    return bad ? 1 : 0;
}
