// lev_band_info.cuh -- the per-pair band of the general-cost kernels, host/device code without CUDA dependencies so
// that the per-pair cores can be compiled for the host and pinned to the oracle there (tests/cpp/diag16_host.cpp).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define TA_BI_HD __host__ __device__ __forceinline__
#else
#define TA_BI_HD inline
#endif

// Per-pair band: the reference's max_k / unit_k clamps (src/levenshtein.rs:400-430) and the Ukkonen band
// [dlo, dlo + W - 1] = [-e, diff + e]; one extra diagonal each side when transpositions need the neighbours' match
// flags.  Shared by every kernel of this file and by the traceback walker so that they agree cell by cell.
TA_BI_HD uint32_t bi_min(uint32_t x, uint32_t y) { return x < y ? x : y; }
struct BandInfo {
    uint32_t max_k;
    int dlo, W;
    bool none;  // the length difference alone exceeds the band: Option::None
};
TA_BI_HD BandInfo band_info(int m, int n, uint32_t k, uint32_t mism, uint32_t gap, uint32_t sgap,
                                              bool trans) {
    BandInfo bi;
    const uint32_t diff = (uint32_t)(n - m);
    uint32_t max_k = bi_min((uint32_t)m * mism, ((uint32_t)m << 1) * gap + (m == 0 ? 0u : sgap + (n == m ? sgap : 0u)));
    max_k = bi_min(k, max_k + diff * gap + (n == m ? 0u : sgap));
    const uint32_t unit_k = (max_k > sgap ? max_k - sgap : 0u) / gap;
    bi.max_k = max_k;
    bi.none = diff > unit_k;
    const uint32_t spare = max_k >= 2 * sgap + diff * gap ? max_k - 2 * sgap - diff * gap : 0u;
    const int e = (int)(spare / (2 * gap)) + (trans ? 1 : 0);
    bi.dlo = -e;
    bi.W = (int)diff + 2 * e + 1;
    return bi;
}


