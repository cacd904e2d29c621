#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include "../triple_accel_b200/csrc/lev_bitpar_core.cuh"
extern "C" {
#include "../oracle/ta_oracle.h"
}
int main(int argc, char** argv) {
    std::mt19937_64 rng(12345);
    std::vector<uint8_t> arena(1 << 16);
    uint8_t* base = (uint8_t*)(((uintptr_t)arena.data() + 4096) & ~(uintptr_t)15);
    long bad = 0, tests = 0;
    for (int it = 0; it < 400000; it++) {
        int alpha = (int[]){2, 3, 4, 26, 256}[rng() % 5];
        int la = rng() % (it % 7 == 0 ? 200 : 40), lb;
        int mode = rng() % 3;
        size_t offa = rng() % 64, offb = 1024 + rng() % 64;
        uint8_t* a = base + offa; uint8_t* b = base + offb;
        for (int i = 0; i < 2048; i++) base[i] = rng() & 0xff;  // junk around
        for (int i = 0; i < la; i++) a[i] = rng() % alpha;
        if (mode == 0) { lb = rng() % (it % 7 == 0 ? 200 : 40); for (int i = 0; i < lb; i++) b[i] = rng() % alpha; }
        else { // mutate
            std::vector<uint8_t> s(a, a + la);
            int ne = rng() % (it % 5 == 0 ? 70 : 12);
            for (int e = 0; e < ne; e++) {
                int kind = rng() % 4;
                if (kind == 0 && !s.empty()) s[rng() % s.size()] = rng() % alpha;
                else if (kind == 1) s.insert(s.begin() + rng() % (s.size() + 1), rng() % alpha);
                else if (kind == 2 && !s.empty()) s.erase(s.begin() + rng() % s.size());
                else if (kind == 3 && s.size() > 1) { size_t p = rng() % (s.size() - 1); std::swap(s[p], s[p + 1]); }
            }
            lb = s.size(); memcpy(b, s.data(), lb);
        }
        for (int trans = 0; trans < 2; trans++) {
            uint32_t kmax = trans ? 30 : 31;
            uint32_t k = rng() % (kmax + 1);
            orc_costs c = {1, 1, 0, (uint8_t)trans};
            uint32_t want = orc_levenshtein_naive_k_with_opts(a, la, b, lb, k, c, NULL, NULL);
            static uint32_t tab[128];
            uint32_t got2 = (it & 1) ? (trans ? bitpar::pair_unit_costs_tab<true,1,uint32_t>(a, la, b, lb, k, (uint8_t*)tab, 4) : bitpar::pair_unit_costs_tab<false,1,uint32_t>(a, la, b, lb, k, (uint8_t*)tab, 4)) : (trans ? bitpar::pair_unit_costs_tab<true,2,uint32_t>(a, la, b, lb, k, (uint8_t*)tab, 4) : bitpar::pair_unit_costs_tab<false,2,uint32_t>(a, la, b, lb, k, (uint8_t*)tab, 4));
            { static uint16_t tab16[128]; uint32_t k16 = rng() % (trans ? 15 : 16); uint32_t want16 = orc_levenshtein_naive_k_with_opts(a, la, b, lb, k16, c, NULL, NULL);
              uint32_t got16 = trans ? bitpar::pair_unit_costs_tab<true,1,uint16_t>(a, la, b, lb, k16, (uint8_t*)tab16, 2) : bitpar::pair_unit_costs_tab<false,1,uint16_t>(a, la, b, lb, k16, (uint8_t*)tab16, 2);
              for (int q = 0; q < 128; q++) if (tab16[q]) { printf("table16 not clean\n"); bad++; tab16[q] = 0; }
              if (got16 != want16) { if (bad++ < 10) printf("TAB16 MISMATCH trans=%d k=%u la=%d lb=%d want=%u got=%u\n", trans, k16, la, lb, want16, got16); } }
            { static uint64_t tab64[64]; uint32_t k64 = rng() % (trans ? 63 : 64); uint32_t want64 = orc_levenshtein_naive_k_with_opts(a, la, b, lb, k64, c, NULL, NULL);
              uint32_t got64 = trans ? bitpar::pair_unit_costs_tab<true,2,uint64_t>(a, la, b, lb, k64, (uint8_t*)tab64, 8) : bitpar::pair_unit_costs_tab<false,2,uint64_t>(a, la, b, lb, k64, (uint8_t*)tab64, 8);
              for (int q = 0; q < 64; q++) if (tab64[q]) { printf("table64 not clean\n"); bad++; tab64[q] = 0; }
              if (got64 != want64) { if (bad++ < 10) printf("TAB64 MISMATCH trans=%d k=%u la=%d lb=%d want=%u got=%u\n", trans, k64, la, lb, want64, got64); } }
            for (int q = 0; q < 128; q++) if (tab[q]) { printf("table not clean\n"); bad++; tab[q] = 0; }
            if (got2 != want) { if (bad++ < 10) printf("TAB MISMATCH trans=%d k=%u la=%d lb=%d want=%u got=%u\n", trans, k, la, lb, want, got2); }
            uint32_t got = trans ? bitpar::pair_unit_costs<true>(a, la, b, lb, k) : bitpar::pair_unit_costs<false>(a, la, b, lb, k);
            tests++;
            if (want != got) { if (bad++ < 10) printf("MISMATCH trans=%d k=%u la=%d lb=%d want=%u got=%u\n", trans, k, la, lb, want, got); }
        }
    }
    printf("tests %ld bad %ld\n", tests, bad);
    return bad != 0;
}
