// lev_diag16_core.cuh -- the per-pair routine of lev_diag16.cu (general EditCosts, one thread per pair, packed u16x2
// anti-diagonals in registers).  Host/device code: on the device the helpers below are single sm_100a instructions
// (VIADDMNMX.U16x2, VIMNMX3.U16x2, VIMNMX.U16x2, VIADD.16x2, PRMT, SHF); on the host they are emulated so that the
// routine can be pinned to the oracle without a GPU (tests/cpp/diag16_host.cpp, tests/test_diag16_host.py).
// The algorithm is described at the top of lev_diag16.cu.
#pragma once

#include <stdint.h>

#include "lev_band_info.cuh"

#if defined(__CUDACC__)
#define D16_HD __host__ __device__ __forceinline__
#else
#define D16_HD inline
#endif

namespace diag16 {

constexpr uint32_t CAP2 = 0x7FFF7FFFu;
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t TOO_WIDE = 0xFFFFFFFEu;  // pair<NR> was asked for a band of more than 4 NR diagonals (never a distance)

D16_HD uint32_t d16_addmin(uint32_t a, uint32_t b, uint32_t c) {  // per half: min(a + b, c)
#if defined(__CUDA_ARCH__)
    return __viaddmin_u16x2(a, b, c);
#else
    const uint32_t lo = ((a & 0xFFFFu) + (b & 0xFFFFu)) & 0xFFFFu, hi = ((a >> 16) + (b >> 16)) & 0xFFFFu;
    const uint32_t cl = c & 0xFFFFu, ch = c >> 16;
    return (lo < cl ? lo : cl) | ((hi < ch ? hi : ch) << 16);
#endif
}
D16_HD uint32_t d16_min2(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __vminu2(a, b);
#else
    const uint32_t al = a & 0xFFFFu, bl = b & 0xFFFFu, ah = a >> 16, bh = b >> 16;
    return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
#endif
}
D16_HD uint32_t d16_min3(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __vimin3_u16x2(a, b, c);
#else
    return d16_min2(d16_min2(a, b), c);
#endif
}
// Packed add of two u16x2 values whose halves cannot carry (every sum in this file stays below 2^16): a plain 32-bit
// add.  On sm_100a that is a VIADD, which issues beside the ALU pipe the kernel is bound by; VIADD.16x2 (__vadd2)
// would take an ALU slot (measured: ncu pipe_alu 93 % with either, 388 M instructions).
D16_HD uint32_t d16_add2(uint32_t a, uint32_t b) { return a + b; }
D16_HD uint32_t d16_fsr(uint32_t lo, uint32_t hi, uint32_t s) {  // low 32 bits of (hi:lo) >> (s & 31)
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    s &= 31u;
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
#endif
}
D16_HD uint32_t d16_fsl(uint32_t lo, uint32_t hi, uint32_t s) {  // high 32 bits of (hi:lo) << (s & 31)
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    s &= 31u;
    return s ? (hi << s) | (lo >> (32u - s)) : hi;
#endif
}
// PTX prmt.b32 (default mode): nibble n of `sel` picks byte (n & 7) of b:a; bit 3 of the nibble replicates that byte's
// sign bit over the result byte instead
D16_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
#else
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t d = 0;
    for (int n = 0; n < 4; n++) {
        const uint32_t nib = (sel >> (4 * n)) & 0xFu;
        uint32_t byte = (uint32_t)(src >> (8 * (nib & 7u))) & 0xFFu;
        if (nib & 8u) byte = (byte & 0x80u) ? 0xFFu : 0u;
        d |= byte << (8 * n);
    }
    return d;
#endif
}
D16_HD uint32_t d16_ldg(const uint32_t *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
// bit 7 of every byte: set iff the bytes of x and y differ
D16_HD uint32_t ne_flags(uint32_t x, uint32_t y) {
    const uint32_t d = x ^ y;
    return (((d | 0x80808080u) - 0x01010101u) | d);
}

// One aligned word of a string per four columns.  next() returns the four bytes at positions pos .. pos + 3 of the
// string (pos may be negative or run past the end: those bytes are arbitrary, the cells that see them are outside the
// matrix) and advances by four.  Words that hold no byte of the string are never loaded.
struct WordStream {
    const uint32_t *base;  // word that holds byte 0 of the string (aligned down)
    int last;              // index of the last word that holds a byte of the string
    int w;                 // index of the word that holds byte `pos`
    uint32_t sh;           // 8 * byte offset of `pos` within its word
    uint32_t cur;          // word w (or 0 when outside the string)
    D16_HD uint32_t load(int idx) const { return (idx >= 0 && idx <= last) ? d16_ldg(base + idx) : 0u; }
    D16_HD void init(const uint8_t *s, int len, int pos) {
        const uint32_t mis = (uint32_t)((uintptr_t)s & 3u);
        base = (const uint32_t *)(s - mis);
        last = ((int)mis + len - 1) >> 2;
        const int p = (int)mis + pos;
        w = p >> 2;  // arithmetic: floor for negative positions
        sh = ((uint32_t)p & 3u) * 8u;
        cur = load(w);
    }
    D16_HD uint32_t next() {
        const uint32_t nx = load(w + 1);
        const uint32_t r = d16_fsr(cur, nx, sh);
        cur = nx;
        w++;
        return r;
    }
};

// levenshtein_naive_k_with_opts(a, b, k, false, costs) for one pair (reference src/levenshtein.rs:376-545): the distance
// if it is <= k, else NONE.  Requires band_info(..).W <= 4 NR and max_k < 0x7F00 (the dispatcher checks both).
template <int NR, bool AFFINE, bool TRANS>
D16_HD uint32_t pair(const uint8_t *pa, uint64_t la, const uint8_t *pb, uint64_t lb, uint32_t k, uint32_t mism,
                     uint32_t gap, uint32_t sgap, uint32_t tcost) {
    constexpr int NW = (2 * NR + 3) / 4;  // 32-bit registers per byte window (2 NR cells per anti-diagonal)
    constexpr int NB = 4 * NW;            // bytes per window (>= cells)
    const bool swap = la > lb;  // src/levenshtein.rs:386 -- "a" is the shorter string
    const uint8_t *ga = swap ? pb : pa;
    const uint8_t *gb = swap ? pa : pb;
    const int m = (int)(swap ? lb : la);
    const int n = (int)(swap ? la : lb);
    const BandInfo bi = band_info(m, n, k, mism, gap, sgap, TRANS);
    if (bi.none) return NONE;  // src/levenshtein.rs:428-430
    if (m == 0) {              // D(0, n) = n*gap + start_gap
        const uint32_t d = (uint32_t)n * gap + (n ? sgap : 0u);
        return d <= bi.max_k ? d : NONE;
    }
    if (bi.W > 4 * NR) return TOO_WIDE;  // the caller re-runs the pair with more registers (lev_diag16.cu)
    const int dlo = bi.dlo;  // <= 0; the host guarantees bi.W <= 4 NR and max_k < 0x7F00
    const uint32_t MM = mism * 0x10001u, GG = gap * 0x10001u, OO = (sgap + gap) * 0x10001u, TT = tcost * 0x10001u;

    // ---- state ---------------------------------------------------------------------------------------------------
    uint32_t E[NR], O[NR];        // D of the even / odd cells (their own values two anti-diagonals ago feed the diagonal move)
    uint32_t EH[NR], EV[NR], OH[NR], OV[NR];  // AFFINE: gap offers to the right / lower neighbour
    uint32_t E4[NR], O4[NR];      // TRANS: D four anti-diagonals ago
    uint32_t FE[NR], FO[NR];      // TRANS: 0xFFFF where the cell's characters differ (previous step of that parity)
    uint32_t Aw[NW], Bw[NW];
#pragma unroll
    for (int r = 0; r < NR; r++) {
        E[r] = O[r] = CAP2;
        EH[r] = EV[r] = OH[r] = OV[r] = CAP2;
        E4[r] = O4[r] = CAP2;
        FE[r] = FO[r] = 0xFFFFFFFFu;
    }
    // E steps run on anti-diagonals s = sE0 + 2t, O steps on s + 1; cell q of an E step: i = iE0 + t - q, j = jE0 + t + q
    const int odd = (-dlo) & 1;
    const int sE0 = -odd;
    const int iE0 = (sE0 - dlo) / 2, jE0 = (sE0 + dlo) / 2;  // exact divisions
    const int s_end = m + n;
    const int t_last = (s_end - sE0) >> 1;  // the step pair that produces anti-diagonal m + n
    // D(0,0) = 0 sits on anti-diagonal 0 = the E step (dlo even) or the O step (dlo odd) of t = 0, at cell q0
    const int q0 = (-dlo) >> 1;
    {
        const uint32_t keep = (q0 & 1) ? 0x0000FFFFu : 0xFFFF0000u, put = OO & ~keep;
#pragma unroll
        for (int r = 0; r < NR; r++) {  // select masks, not indexed stores: the arrays must stay in registers
            const bool hit = r == (q0 >> 1);
            const uint32_t kE = (hit && !odd) ? keep : 0xFFFFFFFFu, pE = (hit && !odd) ? put : 0u;
            const uint32_t kO = (hit && odd) ? keep : 0xFFFFFFFFu, pO = (hit && odd) ? put : 0u;
            E[r] &= kE, O[r] &= kO;
            if (AFFINE) {
                EH[r] = (EH[r] & kE) | pE, EV[r] = (EV[r] & kE) | pE;
                OH[r] = (OH[r] & kO) | pO, OV[r] = (OV[r] & kO) | pO;
            }
        }
    }
    // byte windows as of "after the a-shift of E step t = 0": Aw byte q = a[iE0 - 1 - q], Bw byte q = b[jE0 + q - 1]
#pragma unroll
    for (int x = 0; x < NW; x++) {
        uint32_t wa = 0, wb = 0;
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const int q = 4 * x + y;
            const int ia = iE0 - 1 - q, jb = jE0 + q - 1;
            wa |= (uint32_t)((ia >= 0 && ia < m) ? ga[ia] : 0) << (8 * y);
            wb |= (uint32_t)((jb >= 0 && jb < n) ? gb[jb] : 0) << (8 * y);
        }
        Aw[x] = wa, Bw[x] = wb;
    }

    // 0x00FF per half where the two cells of register r differ (bit 7 of the flag bytes replicated over the low byte)
    auto expand = [&](const uint32_t *F, int r) -> uint32_t {
        return prmt(F[r >> 1], 0u, (r & 1) ? 0x4B4Au : 0x4948u);
    };
    // 0xFFFF per half where they differ (TRANS: the neighbours' flags gate the transposition)
    auto expand_full = [&](const uint32_t *F, int r) -> uint32_t {
        return prmt(F[r >> 1], 0u, (r & 1) ? 0xBBAAu : 0x9988u);
    };
    // one cell pair: returns D; updates the offers
    auto cell = [&](uint32_t diag, uint32_t ne16, uint32_t h, uint32_t v, uint32_t &offH, uint32_t &offV, uint32_t d4,
                    uint32_t both_ne) -> uint32_t {
        // diag + (differ ? mismatch : 0) = min(diag + 255 * differ, diag + mismatch): the plain add runs beside the ALU pipe
        const uint32_t t1 = d16_addmin(diag, ne16, d16_add2(diag, MM));
        uint32_t d;
        if (AFFINE) {
            d = d16_min3(t1, h, v);  // h, v are the neighbours' offers
        } else {
            d = d16_addmin(h, GG, t1);  // h, v are the neighbours' D
            d = d16_addmin(v, GG, d);
        }
        if (TRANS) {
            const uint32_t tr = d16_add2(d4, TT) | both_ne;  // 0xFFFF where no transposition ends here
            d = d16_min2(d, tr);
        }
        if (AFFINE) {
            const uint32_t t = d16_add2(d, OO);
            offH = d16_addmin(h, GG, t);
            offV = d16_addmin(v, GG, t);
        }
        return d;
    };
    // E step: `top` carries the next byte of a in its top byte
    auto stepE = [&](uint32_t top) {
#pragma unroll
        for (int x = NW - 1; x > 0; x--) Aw[x] = d16_fsl(Aw[x - 1], Aw[x], 8);
        Aw[0] = d16_fsl(top, Aw[0], 8);
        uint32_t F[NW];
#pragma unroll
        for (int x = 0; x < NW; x++) F[x] = ne_flags(Aw[x], Bw[x]);
        uint32_t nf[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const uint32_t ne16 = expand(F, r);
            if (TRANS) nf[r] = expand_full(F, r);
            // left neighbours = odd cells (2r - 1, 2r), upper neighbours = odd cells (2r, 2r + 1)
            const uint32_t hl = r ? (AFFINE ? OH[r - 1] : O[r - 1]) : CAP2;
            const uint32_t h = prmt(hl, AFFINE ? OH[r] : O[r], 0x5432u);
            const uint32_t v = AFFINE ? OV[r] : O[r];
            uint32_t both = 0;
            if (TRANS) both = prmt(r ? FO[r - 1] : 0xFFFFFFFFu, FO[r], 0x5432u) | FO[r];
            const uint32_t old = E[r];
            E[r] = cell(old, ne16, h, v, EH[r], EV[r], E4[r], both);
            if (TRANS) E4[r] = old;
        }
        if (TRANS) {
#pragma unroll
            for (int r = 0; r < NR; r++) FE[r] = nf[r];
        }
    };
    // O step: `low` carries the next byte of b in its low byte
    auto stepO = [&](uint32_t low) {
#pragma unroll
        for (int x = 0; x < NW - 1; x++) Bw[x] = d16_fsr(Bw[x], Bw[x + 1], 8);
        Bw[NW - 1] = d16_fsr(Bw[NW - 1], low, 8);
        uint32_t F[NW];
#pragma unroll
        for (int x = 0; x < NW; x++) F[x] = ne_flags(Aw[x], Bw[x]);
        uint32_t nf[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const uint32_t ne16 = expand(F, r);
            if (TRANS) nf[r] = expand_full(F, r);
            // left neighbours = even cells (2r, 2r + 1), upper neighbours = even cells (2r + 1, 2r + 2)
            const uint32_t h = AFFINE ? EH[r] : E[r];
            const uint32_t vu = (r + 1 < NR) ? (AFFINE ? EV[r + 1] : E[r + 1]) : CAP2;
            const uint32_t v = prmt(AFFINE ? EV[r] : E[r], vu, 0x5432u);
            uint32_t both = 0;
            if (TRANS) both = FE[r] | prmt(FE[r], (r + 1 < NR) ? FE[r + 1] : 0xFFFFFFFFu, 0x5432u);
            const uint32_t old = O[r];
            O[r] = cell(old, ne16, h, v, OH[r], OV[r], O4[r], both);
            if (TRANS) O4[r] = old;
        }
        if (TRANS) {
#pragma unroll
            for (int r = 0; r < NR; r++) FO[r] = nf[r];
        }
    };

    // Cells grow by at most 2 * 255 per column between clamps (cells <= CAP + 16 * 510, offers 510 more: < 2^16)
    auto clamp_all = [&]() {
#pragma unroll
        for (int r = 0; r < NR; r++) {
            E[r] = d16_min2(E[r], CAP2), O[r] = d16_min2(O[r], CAP2);
            if (AFFINE) {
                EH[r] = d16_min2(EH[r], CAP2), EV[r] = d16_min2(EV[r], CAP2);
                OH[r] = d16_min2(OH[r], CAP2), OV[r] = d16_min2(OV[r], CAP2);
            }
            if (TRANS) E4[r] = d16_min2(E4[r], CAP2), O4[r] = d16_min2(O4[r], CAP2);
        }
    };

    // ---- t = 0: the E step has nothing to compute (its only cell inside the matrix, if any, is the planted D(0,0));
    //      the O step runs when anti-diagonal 0 was the E step, else D(0,0) was planted into O and it is skipped too
    WordStream sa, sb;
    {
        const int jb = jE0 + NB - 1;
        const uint32_t low = (jb >= 0 && jb < n) ? gb[jb] : 0u;
        if (!odd) {
            stepO(low);
        } else {
#pragma unroll
            for (int x = 0; x < NW - 1; x++) Bw[x] = d16_fsr(Bw[x], Bw[x + 1], 8);
            Bw[NW - 1] = d16_fsr(Bw[NW - 1], low, 8);
        }
    }
    sa.init(ga, m, iE0);          // E step t >= 1 shifts in a[iE0 + t - 1]
    sb.init(gb, n, jE0 + NB);     // O step t >= 1 shifts in b[jE0 + t + NB - 1]

    // ---- main loop: four columns per iteration --------------------------------------------------------------------
    int t = 1;
    for (; t + 3 <= t_last; t += 4) {
        const uint32_t wa = sa.next(), wb = sb.next();
        stepE(wa << 24);
        stepO(wb);
        stepE(wa << 16);
        stepO(wb >> 8);
        stepE(wa << 8);
        stepO(wb >> 16);
        stepE(wa);
        stepO(wb >> 24);
        if ((t & 12) == 12) clamp_all();  // every 16 columns (t = 13, 29, ...): see "why clamping is exact" in lev_diag16.cu
    }
    if (t <= t_last) {  // up to three more columns
        const uint32_t wa = sa.next(), wb = sb.next();
#pragma unroll 1
        for (int u = 0; t <= t_last; t++, u++) {
            stepE(wa << (24 - 8 * u));
            stepO(wb >> (8 * u));
        }
    }

    // ---- the cell (m, n): anti-diagonal m + n, diagonal n - m -------------------------------------------------------
    const int pf = (s_end - dlo) & 1;
    const int qf = ((n - m) - dlo - pf) >> 1;
    uint32_t val = 0xFFFFu;
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const uint32_t x = pf ? O[r] : E[r];
        if (r == (qf >> 1)) val = (qf & 1) ? (x >> 16) : (x & 0xFFFFu);
    }
    return val <= bi.max_k ? val : NONE;
}

}  // namespace diag16
