// lev_bitpar_core.cuh -- per-pair core of the bit-parallel banded kernel (unit costs), host/device.
//
// One thread owns one pair.  The k-band (Ukkonen: W = diff + 2e + 1 <= 32 diagonals, e = (max_k - diff)/2) is a
// 32-row window that slides one row down per text column; the vertical deltas of the window's cells live in two
// 32-bit words (VP/VN) and a column is advanced with Myers' carry trick in Hyyro's diagonal-aligned form
//     D0 = (((Eq & VP) + VP) ^ VP) | Eq | VN [| TR]        X  = D0 >> 1
//     HP = VN | ~(D0 | VP),  HN = D0 & VP                  VP' = HN | ~(X | HP),  VN' = X & HP
// (TR = ~D0prev & (Eq << 1) & (Eqprev >> 1) adds restricted-Damerau transpositions).  Rows above the matrix are
// "virtual" cells with D = j - i (vertical delta -1), which makes the first row D[0][j] = j come out of the same
// recurrence; rows below the pattern are garbage that can never flow upwards.  The distance is read off the
// diagonal through (m, n):  d = diff + n - #columns whose D0 bit at that diagonal is set.
//
// Eq (which of the 32 window rows match the column's byte) is a SIMD-in-register compare: the window's pattern
// bytes sit transposed in 8 registers (register q, byte t <-> row q + 8t), each is XORed with the broadcast text
// byte and zero bytes are detected exactly with ((x & 0x7f..) + 0x7f..) | x; shifting register q's flags right by
// 7 - q and OR-ing the eight results puts row r's flag at bit r.  Sliding the window is one PRMT (drop byte 0 of
// the oldest register, append the next pattern byte); with the column loop unrolled by 16 every register role,
// byte selector and shift is a compile-time constant.
//
// Both strings are streamed straight from global memory in aligned 16-byte vectors (one LDG.128 per string per 16
// columns, prefetched one iteration ahead) and re-aligned in registers (a two-stage word select plus a funnel
// shift), so arbitrary CSR offsets and string lengths need no staging buffer.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TA_HD __host__ __device__ __forceinline__
#else
#define TA_HD inline
#endif

namespace bitpar {

struct V4 {
    uint32_t x, y, z, w;
};

TA_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {  // low 32 bits of (hi:lo) >> s, s in [0, 31]
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}

TA_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    uint32_t r;  // inline PTX: __byte_perm() ignores bit 3 of a selector nibble (the sign-replicate mode used below)
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
#else
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t nib = (sel >> (4 * i)) & 15u;
        uint32_t byte = (uint32_t)((v >> (8 * (nib & 7))) & 0xff);
        if (nib & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;  // default mode: bit 3 replicates the byte's sign bit
        r |= byte << (8 * i);
    }
    return r;
#endif
}

TA_HD V4 ldvec(const uint8_t *p) {  // p is 16-byte aligned
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg((const uint4 *)p);
    return V4{v.x, v.y, v.z, v.w};
#else
    V4 v;
    const uint32_t *q = (const uint32_t *)p;
    v.x = q[0], v.y = q[1], v.z = q[2], v.w = q[3];
    return v;
#endif
}

// A byte stream read through aligned 16-byte vectors.  Logical byte t of the stream is memory byte start + t;
// vector addresses are clamped into [lo, hi] (the vectors holding the string's first / last byte), so nothing
// outside the string's own 16-byte-aligned extent is ever touched; clamped positions deliver don't-care bytes.
struct Stream {
    const uint8_t *base;  // 16-byte aligned address of vector 0
    int next;             // index of the next vector to fetch (unclamped)
    int lo, hi;           // clamp range of vector indices
    uint32_t wsel;        // word offset of the stream start inside its first vector (0..3)
    uint32_t bsh;         // 8 * byte offset inside the word (0, 8, 16, 24)
    V4 cur, nxt;          // vectors c and c+1 of the stream

    TA_HD V4 fetch() {
        int i = next < hi ? next : hi;
        i = i > lo ? i : lo;
        next += 1;
        return ldvec(base + (intptr_t)i * 16);
    }
    // start: address of logical byte 0 (may lie before first); first/last: first and last valid byte addresses
    TA_HD void init(intptr_t start, uintptr_t first, uintptr_t last) {
        const uintptr_t s = (uintptr_t)start;
        const uintptr_t b0 = s & ~(uintptr_t)15;
        base = (const uint8_t *)b0;
        lo = (int)(((intptr_t)(first & ~(uintptr_t)15) - (intptr_t)b0) >> 4);
        hi = (int)(((intptr_t)(last & ~(uintptr_t)15) - (intptr_t)b0) >> 4);
        next = 0;
        wsel = (uint32_t)(s >> 2) & 3u;
        bsh = ((uint32_t)s & 3u) * 8u;
        cur = fetch();
        nxt = fetch();
    }
    // the next 16 logical bytes as four little-endian words; advances the stream
    TA_HD void take(uint32_t out[4]) {
        if ((wsel | bsh) == 0) {  // 16-byte aligned stream: the vector is the chunk
            out[0] = cur.x, out[1] = cur.y, out[2] = cur.z, out[3] = cur.w;
        } else {
            const uint32_t v[8] = {cur.x, cur.y, cur.z, cur.w, nxt.x, nxt.y, nxt.z, nxt.w};
            uint32_t z[7], y[5];
            const bool s1 = wsel & 1u, s2 = wsel & 2u;
#pragma unroll
            for (int i = 0; i < 7; i++) z[i] = s1 ? v[i + 1] : v[i];
#pragma unroll
            for (int i = 0; i < 5; i++) y[i] = s2 ? z[i + 2] : z[i];
#pragma unroll
            for (int i = 0; i < 4; i++) out[i] = funnel_r(y[i], y[i + 1], bsh);
        }
        cur = nxt;
        nxt = fetch();
    }
};

// flags (bit 7 of each byte) of the bytes of r that equal the corresponding byte of bb
TA_HD uint32_t eq_flags(uint32_t r, uint32_t bb) {
    const uint32_t x = r ^ bb;
    const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;  // bit 7 of each byte: low 7 bits non-zero
    return ~(t | x) & 0x80808080u;
}

// Distance of one pair, `a` being the shorter string (m <= n), unit costs.  Requires m >= 1 and
// diff + 2e + 1 <= 32 with e = (max_k - diff)/2 (+1 if TRANS and max_k - diff is odd).  Returns the exact distance whenever it is
// <= max_k (and some value > max_k otherwise).
template <bool TRANS>
TA_HD uint32_t distance32(const uint8_t *a, int m, const uint8_t *b, int n, uint32_t max_k) {
    const int diff = n - m;
    // one margin diagonal each side so that the transposition test can see its neighbours' match flags -- only needed
    // when max_k - diff is odd: with an even budget the gaps needed to reach an extreme diagonal and come back use
    // all of max_k, so no transposition can lie on it
    const int e = (int)((max_k - (uint32_t)diff) >> 1) + ((TRANS && ((max_k - (uint32_t)diff) & 1u)) ? 1 : 0);
    const int dhi = diff + e;  // window row p of column j is matrix row i = j - dhi + p

    Stream sa, sb;
    sa.init((intptr_t)a - dhi, (uintptr_t)a, (uintptr_t)a + m - 1);
    sb.init((intptr_t)b, (uintptr_t)b, (uintptr_t)b + n - 1);

    // initial window = pattern-stream bytes 0..31, transposed: register q, byte t <- stream byte q + 8t
    uint32_t R[8];
    {
        uint32_t x[8];
        sa.take(x);
        sa.take(x + 4);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int w = q >> 2, bsel = q & 3;
            const uint32_t lo2 = prmt(x[w], x[w + 2], (uint32_t)(bsel | ((4 + bsel) << 4)));      // bytes t = 0, 1
            const uint32_t hi2 = prmt(x[w + 4], x[w + 6], (uint32_t)(bsel | ((4 + bsel) << 4)));  // bytes t = 2, 3
            R[q] = prmt(lo2, hi2, 0x5410);
        }
    }

    uint32_t VP = dhi >= 32 ? 0u : (0xffffffffu << dhi);  // rows i >= 1: vertical delta +1 in column 0
    uint32_t VN = ~VP;                                    // rows i <= 0: virtual cells, delta -1
    uint32_t D0prev = 0xffffffffu, Eqprev = 0;
    uint32_t matches = 0;  // columns whose final-diagonal cell has diagonal delta 0

    const uint32_t emask = 1u << e;  // the diagonal through (m, n) is window row e
    for (int j0 = 0; j0 < n; j0 += 16) {
        uint32_t aw[4], bw[4];
        sa.take(aw);
        sb.take(bw);
        uint32_t hist = 0;  // bit e + u <- column u's D0 bit on the final diagonal (e <= 16)
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const uint32_t bsel = (uint32_t)(u & 3);
            const uint32_t bb = prmt(bw[u >> 2], 0, bsel * 0x1111u);  // text byte of this column, broadcast
            // role q is played by register (q + u) & 7 (the window slid u rows since the chunk began)
            uint32_t Eq = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) Eq |= eq_flags(R[(q + u) & 7], bb) >> (7 - q);
            uint32_t D0 = (((Eq & VP) + VP) ^ VP) | Eq | VN;
            if (TRANS) {
                D0 |= ~D0prev & (Eq << 1) & (Eqprev >> 1);
                D0prev = D0;
                Eqprev = Eq;
            }
            const uint32_t HP = VN | ~(D0 | VP);
            const uint32_t HN = D0 & VP;
            const uint32_t X = D0 >> 1;
            VN = X & HP;
            VP = HN | ~(X | HP);
            hist += (D0 & emask) << u;
            // slide: the oldest register (role 0) drops its byte 0 and takes the next pattern byte as role 7
            R[u & 7] = prmt(R[u & 7], aw[u >> 2], 0x0321u | ((4u + bsel) << 12));
        }
        const int cols = n - j0;  // columns of this chunk that exist
        const uint32_t valid = cols >= 16 ? 0xffffu : ((1u << cols) - 1u);
#if defined(__CUDA_ARCH__)
        matches += __popc((hist >> e) & valid);
#else
        matches += (uint32_t)__builtin_popcount((hist >> e) & valid);
#endif
    }
    return (uint32_t)diff + (uint32_t)n - matches;
}

// ---------------------------------------------------------------------------------------------------------------
// distance32_tab: same recurrence, but Eq comes from a per-thread match table instead of SIMD compares.
//
// tab[c * stride] (c = low 7 bits of a byte, 128 entries, 32 bits each, thread-private column of a shared-memory
// array) has bit (t mod 32) set iff pattern-stream byte t is inside the current 32-row window and its low 7 bits
// are c; the bytes' top bits live in a 32-bit plane A7 with the same circular bit numbering.  Sliding the window
// clears the bit of the byte that leaves and sets the same bit position for the byte that enters (two read-modify-
// writes), so a column costs  Eq = rotr(tab[b & 0x7f] & ~(A7 ^ topmask(b)), u mod 32)  -- one LDS, one LOP3, one
// funnel shift -- instead of 8 SWAR compares.  The table must be all-zero on entry and is left all-zero on exit.
// `tab` points at this thread's entry 0; entry c lives c * pitch BYTES further (pitch = sizeof(W) * threads).  The
// address is one multiply-add (IMAD, FMA pipe) -- a shift + add would cost an extra ALU-pipe slot per access.
template <typename W>
TA_HD W &tab_at(uint8_t *tab, uint32_t c, uint32_t pitch) {
    return *(W *)(tab + c * pitch);
}
TA_HD uint32_t byte_of(uint32_t w, int t) {  // byte t of w, zero-extended (one PRMT)
    return prmt(w, 0u, 0x4440u | (uint32_t)t);
}
// gathers bit `bitpos` of each of the 16 bytes of w[0..3] into a 16-bit word (byte t of w[i] -> bit 4i + t)
TA_HD uint32_t gather_bits16(const uint32_t w[4], int bitpos) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t f = (w[i] >> bitpos) & 0x01010101u;        // bits 0, 8, 16, 24
        r |= (((f * 0x00204081u) >> 21) & 0xfu) << (4 * i);       // gathered to 4 adjacent bits
    }
    return r;
}
// all-ones iff bit `bitpos` of byte t of w is set
template <typename W>
TA_HD W bit_mask_of(uint32_t w, int t, int bitpos) {
    return (W)(int64_t)((int32_t)(w << (31 - 8 * t - bitpos)) >> 31);
}
template <typename W>
TA_HD W rotr(W x, uint32_t s) {  // s in [0, bits)
    if (sizeof(W) == 4) return (W)funnel_r((uint32_t)x, (uint32_t)x, s);
    if (sizeof(W) == 2) {  // duplicate into both halves, then one 32-bit shift
        const uint32_t d = prmt((uint32_t)x, 0u, 0x1010u);
        return (W)(d >> s);
    }
    return s ? (W)((x >> s) | (x << (8 * sizeof(W) - s))) : x;
}
template <typename W>
struct WiderOf {
    typedef uint32_t type;
};
template <>
struct WiderOf<uint64_t> {
    typedef uint64_t type;
};
TA_HD uint32_t popc_w(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// distance_tab: same recurrence as distance32, but Eq comes from a per-thread match table instead of SIMD compares.
//
// tab[c] (c = the byte's low class bits; thread-private column of a shared-memory array) has bit (t mod BITS) set
// iff pattern-stream byte t is inside the current BITS-row window and its class is c; the bytes' remaining top
// bits live in register planes with the same circular bit numbering.  Sliding the window clears the bit of the
// byte that leaves and sets the same bit position for the byte that enters (two read-modify-writes), so a column
// costs  Eq = rotr(tab[class(b)] & ~miss, u mod BITS)  -- one LDS, two or three LOP3, one rotate -- instead of 8
// SWAR compares.  The table must be all-zero on entry and is left all-zero on exit.
//   W = uint16_t: 16-row window (k <= 15 / 14), half the shared memory per thread;  W = uint32_t: 32-row window
//   (k <= 31; <= 30 with transpositions);  W = uint64_t: 64-row window (k <= 63 / 62).
//   PLANES = 1: 128 entries (7-bit classes) + the A7 plane;  PLANES = 2: 64 entries (6-bit classes) + A6, A7.
template <bool TRANS, int PLANES, typename W>
TA_HD uint32_t distance_tab(const uint8_t *a, int m, const uint8_t *b, int n, uint32_t max_k, uint8_t *tab,
                            const uint32_t pitch) {
    constexpr int BITS = 8 * (int)sizeof(W);
    constexpr int RING = BITS / 4;  // words holding the window's bytes
    constexpr uint32_t CMASK = PLANES == 1 ? 0x7f7f7f7fu : 0x3f3f3f3fu;
    typedef typename WiderOf<W>::type HistT;  // per-chunk match history: bit e + u, e <= BITS/2
    const int diff = n - m;
    // one margin diagonal each side so that the transposition test can see its neighbours' match flags -- only needed
    // when max_k - diff is odd: with an even budget the gaps needed to reach an extreme diagonal and come back use
    // all of max_k, so no transposition can lie on it
    const int e = (int)((max_k - (uint32_t)diff) >> 1) + ((TRANS && ((max_k - (uint32_t)diff) & 1u)) ? 1 : 0);
    const int dhi = diff + e;  // window row p of column j is matrix row i = j - dhi + p

    Stream sa, sb;
    sa.init((intptr_t)a - dhi, (uintptr_t)a, (uintptr_t)a + m - 1);
    sb.init((intptr_t)b, (uintptr_t)b, (uintptr_t)b + n - 1);

    // ring of the pattern-stream bytes currently inside the window (class bits only), oldest chunk first
    uint32_t ring[RING];
    W A7 = 0, A6 = 0;
    {
        uint32_t x[RING];
#pragma unroll
        for (int c = 0; c < RING / 4; c++) sa.take(x + 4 * c);
#pragma unroll
        for (int c = 0; c < RING / 4; c++) {
            A7 |= (W)((W)gather_bits16(x + 4 * c, 7) << (16 * c));
            if (PLANES == 2) A6 |= (W)((W)gather_bits16(x + 4 * c, 6) << (16 * c));
        }
#pragma unroll
        for (int w = 0; w < RING; w++) ring[w] = x[w] & CMASK;
#pragma unroll
        for (int t = 0; t < BITS; t++) tab_at<W>(tab, byte_of(ring[t >> 2], t & 3), pitch) |= (W)((W)1 << t);
    }

    const W ones = (W) ~(W)0;
    W VP = dhi >= BITS ? (W)0 : (W)(ones << dhi);
    W VN = (W)~VP;
    W D0prev = ones, Eqprev = 0;
    uint32_t matches = 0;
    const W emask = (W)((W)1 << e);
    HistT acc = 0;       // (# columns of this chunk whose final-diagonal cell has diagonal delta 0) << e
    uint32_t phase = 0;  // circular bit position of this chunk's first column, (16 * chunk) mod BITS
    W bit0 = 1;          // 1 << phase
    W tops7 = 0, tops6 = 0;  // plane bits of the chunk's entering bytes at their circular positions

    // one DP column: b7m/b6m = all-ones iff the text byte has bit 7/6 set, bcls = its class bits, lv/en = classes of
    // the pattern bytes that leave/enter the window after this column, bit = their circular position, rot = the
    // circular position of window row 0
    // (a7v/a6v = the planes as of this column; upd = false when the caller derives them itself)
    auto column = [&](W b7m, W b6m, uint32_t bcls, uint32_t lv, uint32_t en, W bit, uint32_t rot, W a7v, W a6v,
                      const bool upd) {
        W miss = (W)(a7v ^ b7m);  // rows whose plane bits differ from the text byte's
        if (PLANES == 2) miss |= (W)(a6v ^ b6m);
        const W raw = (W)(tab_at<W>(tab, bcls, pitch) & ~miss);
        const W Eq = rotr<W>(raw, rot);
        W D0 = (W)((((Eq & VP) + VP) ^ VP) | Eq | VN);
        if (TRANS) {
            D0 |= (W)(~D0prev & (Eq << 1) & (Eqprev >> 1));
            D0prev = D0;
            Eqprev = Eq;
        }
        const W HP = (W)(VN | ~(D0 | VP));
        const W HN = (W)(D0 & VP);
        const W X = (W)(D0 >> 1);
        VN = (W)(X & HP);
        VP = (W)(HN | ~(X | HP));
        acc += (HistT)(D0 & emask);
        tab_at<W>(tab, lv, pitch) &= (W)~bit;
        tab_at<W>(tab, en, pitch) |= bit;
        if (upd) {
            A7 = (W)((A7 & ~bit) | (tops7 & bit));
            if (PLANES == 2) A6 = (W)((A6 & ~bit) | (tops6 & bit));
        }
    };

    uint32_t aw[4] = {0, 0, 0, 0}, bw[4], bc[4];
    int lim = n;
    int j0 = 0;
    for (; j0 + 16 <= lim; j0 += 16) {  // full 16-column chunks: every selector and shift below is a constant
        const uint32_t m_old = matches;
        sa.take(aw);                  // the bytes that enter during this chunk
        sb.take(bw);
        tops7 = (W)((W)gather_bits16(aw, 7) * bit0);
        if (PLANES == 2) tops6 = (W)((W)gather_bits16(aw, 6) * bit0);
#pragma unroll
        for (int w = 0; w < 4; w++) {
            aw[w] &= CMASK;
            bc[w] = bw[w] & CMASK;
        }
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (BITS == 16) {
                // 16-row window: a chunk is one full turn of the circular numbering, so the planes as of column u
                // are a constant-mask select between the chunk-start planes and the entering bytes' planes
                const W lm = (W)((1u << u) - 1u);
                column(bit_mask_of<W>(bw[u >> 2], u & 3, 7), PLANES == 2 ? bit_mask_of<W>(bw[u >> 2], u & 3, 6) : (W)0,
                       byte_of(bc[u >> 2], u & 3), byte_of(ring[u >> 2], u & 3), byte_of(aw[u >> 2], u & 3),
                       (W)(1u << u), (uint32_t)u, (W)((A7 & ~lm) | (tops7 & lm)),
                       PLANES == 2 ? (W)((A6 & ~lm) | (tops6 & lm)) : (W)0, false);
            } else {
                column(bit_mask_of<W>(bw[u >> 2], u & 3, 7), PLANES == 2 ? bit_mask_of<W>(bw[u >> 2], u & 3, 6) : (W)0,
                       byte_of(bc[u >> 2], u & 3), byte_of(ring[u >> 2], u & 3), byte_of(aw[u >> 2], u & 3),
                       (W)(bit0 << u), phase + (uint32_t)u, A7, A6, true);
            }
        }
        if (BITS == 16) {
            A7 = tops7;
            A6 = tops6;
        }
        matches += (uint32_t)(acc >> e);
        acc = 0;
#pragma unroll
        for (int w = 0; w + 4 < RING; w++) ring[w] = ring[w + 4];  // drop the oldest chunk, append the new one
#pragma unroll
        for (int w = 0; w < 4; w++) ring[RING - 4 + w] = aw[w];
        phase = (phase + 16u) & (uint32_t)(BITS - 1);
        bit0 = (W)((W)1 << phase);
        // early exit: the final-diagonal value diff + j - matches(j) never decreases and ends as the result (tested
        // on the count as of the start of this chunk, which is long computed: the branch never waits for the chain)
        lim = (uint32_t)diff + (uint32_t)j0 - m_old > max_k ? 0 : lim;  // folded into the loop bound: single exit
    }
    const bool dead = lim == 0;  // (n >= 1 here)
    if (!dead && j0 < n) {  // last n % 16 columns: the same column step, rolled, bytes shifted out of the chunk's words
        sa.take(aw);
        sb.take(bw);
        tops7 = (W)((W)gather_bits16(aw, 7) * bit0);
        if (PLANES == 2) tops6 = (W)((W)gather_bits16(aw, 6) * bit0);
#pragma unroll
        for (int w = 0; w < 4; w++) aw[w] &= CMASK;
        uint32_t ea[4] = {aw[0], aw[1], aw[2], aw[3]};
        uint32_t lv[4] = {ring[0], ring[1], ring[2], ring[3]};
        for (int u = 0; u < n - j0; u++) {
            const uint32_t bch = bw[0] & 0xffu;
            column((W)(0u - (W)(bch >> 7)), (W)(0u - (W)((bch >> 6) & 1u)), bch & (CMASK & 0xffu), lv[0] & 0xffu,
                   ea[0] & 0xffu, (W)(bit0 << u), phase + (uint32_t)u, A7, A6, true);
#pragma unroll
            for (int w = 0; w < 3; w++) {
                bw[w] = funnel_r(bw[w], bw[w + 1], 8);
                ea[w] = funnel_r(ea[w], ea[w + 1], 8);
                lv[w] = funnel_r(lv[w], lv[w + 1], 8);
            }
            bw[3] >>= 8;
            ea[3] >>= 8;
            lv[3] >>= 8;
        }
        matches += (uint32_t)(acc >> e);
    }
    // leave the table clean: every set bit belongs to a byte of the ring or of the last chunk taken
#pragma unroll
    for (int t = 0; t < BITS; t++) tab_at<W>(tab, byte_of(ring[t >> 2], t & 3), pitch) = 0;
#pragma unroll
    for (int t = 0; t < 16; t++) tab_at<W>(tab, byte_of(aw[t >> 2], t & 3), pitch) = 0;
    return dead ? 0xFFFFFFFFu : (uint32_t)diff + (uint32_t)n - matches;
}

// ---------------------------------------------------------------------------------------------------------------
// distance_blk: the match-table recurrence with BLOCK-wise table maintenance (bands of W <= 33 - C diagonals).
//
// Only the band's W rows have to see their exact match flags; every other row of the 32-row window merely has to
// stay an upper bound of the true DP, i.e. its Eq bits may be dropped but never invented.  So the table does not
// have to follow the window exactly.  It is a 32-position circular frame (stream byte t sits at bit t mod 32) cut
// into blocks of C positions (C = 16: the two halves of an entry; C = 8: its four bytes).  With L = 32 - C:
//   * before column j is looked up, stream byte j + L is OR-ed in (one read-modify-write per column, independent of
//     the recurrence, so its latency hides behind it) -- the band's last row of column j is byte j + W - 1 <= j + L;
//   * at the start of each chunk of C columns [g, g + C) the block [g - C, g), which left the window, is ZEROED BY
//     VALUE: a sub-word store of 0 to the entry of each of its bytes -- no read, no mask, because every bit of that
//     sub-word belongs to the same (dead) block.  These are the positions the chunk's entering bytes will use.
// A column's Eq is rotr(tab[class(b)] [& plane], j mod 32).  The rotation wraps bytes that already left the window
// (up to C - 1 of them, not yet zeroed) into rows p >= 33 - C, where they can INVENT matches.  That is harmless: with
// unit costs every change of diagonal costs 1, the path starts on row dhi and ends on row e, so a path through such
// a cell costs at least (p - dhi) + (p - e) >= 66 - 2C - (W - 1) >= 34 - C > max_k (max_k <= W <= 33 - C); it can
// neither lower a distance that is <= max_k nor pull one that is > max_k below the threshold.
// Columns are unrolled 32 deep so that every rotation, bit and sub-word offset is an immediate.  The last n mod 16
// columns slide the table column by column with an exact clear (rolled code).
//   PLANES = 1: 128 entries (7-bit classes) + the A7 plane;  PLANES = 0: 256 entries, no plane.
//   `tab` = the thread's entry 0, `pitch` = bytes between entries; the table is all-zero on entry and on exit.
// The bytes a thread will want first when it starts its NEXT work item: pulled into L1 when the current item leaves
// its main loop (the tail and the table clean-up that follow hide the L2 latency); device only.
struct NextHint {
    const uint8_t *p[4];  // first byte of each string (nullptr: none)
    uint32_t len[4];
};
TA_HD void prefetch_next(const NextHint *h) {
#if defined(__CUDA_ARCH__)
    if (h) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (h->p[i]) {  // the line of the first byte and the line 127 bytes on (a ragged string straddles two)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(h->p[i]));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(h->p[i] + (h->len[i] < 128u ? h->len[i] - 1u : 127u)));
            }
    }
#else
    (void)h;
#endif
}
TA_HD uint8_t *blk_entry(uint8_t *tab, uint32_t w, int t, uint32_t pitch) { return tab + byte_of(w, t) * pitch; }
template <int V>
struct IntC {
    static constexpr int value = V;
};

template <bool TRANS, int PLANES, int C>
TA_HD uint32_t distance_blk(const uint8_t *a, int m, const uint8_t *b, int n, uint32_t max_k, uint8_t *tab,
                            const uint32_t pitch, const NextHint *next = nullptr) {
    static_assert(C == 8 || C == 16, "block = a byte or a half of the entry");
    static_assert(PLANES == 0 || PLANES == 1, "");
    typedef uint32_t W;
    constexpr uint32_t CMASK = PLANES == 1 ? 0x7f7f7f7fu : 0xffffffffu;
    constexpr int L = 32 - C;              // lead of the inserted byte over the column
    constexpr int NT = C == 16 ? 1 : 2;    // 16-byte takes of the pattern stream the table runs ahead of the text
    constexpr int NW = 8 + 4 * NT;         // window words: stream bytes [16 s - 16, 16 s + 16 (NT + 1))
    const int diff = n - m;
    const int e = (int)((max_k - (uint32_t)diff) >> 1) + ((TRANS && ((max_k - (uint32_t)diff) & 1u)) ? 1 : 0);
    const int dhi = diff + e;  // window row p of column j is matrix row i = j - dhi + p

    Stream sa, sb;
    sa.init((intptr_t)a - dhi, (uintptr_t)a, (uintptr_t)a + m - 1);
    sb.init((intptr_t)b, (uintptr_t)b, (uintptr_t)b + n - 1);

    // win = class bits of pattern-stream bytes [16 s - 16, ...) at superstep s; g7 = top bits of the last takes
    uint32_t win[NW];
    uint32_t g7prev = 0, g7cur = 0;
    W A7 = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) win[w] = 0;
    {
        uint32_t x[4 * NT];
#pragma unroll
        for (int c = 0; c < NT; c++) sa.take(x + 4 * c);
        if (PLANES) {
            g7cur = gather_bits16(x + 4 * (NT - 1), 7);
            if (NT == 2) g7prev = gather_bits16(x, 7);
            A7 = NT == 1 ? g7cur : (g7prev | ((g7cur & 0xffu) << 16));
        }
#pragma unroll
        for (int w = 0; w < 4 * NT; w++) win[4 + w] = x[w] & CMASK;
#pragma unroll
        for (int t = 0; t < L; t++) *(W *)blk_entry(tab, win[4 + (t >> 2)], t & 3, pitch) |= 1u << t;
    }

    W VP = dhi >= 32 ? 0u : (0xffffffffu << dhi);
    W VN = ~VP;
    W D0prev = 0xffffffffu, Eqprev = 0;
    uint32_t matches = 0;
    const W emask = 1u << e;
    uint32_t acc = 0;
    uint32_t bw[4], bc[4];

    // the recurrence for one column given its match word
    auto step = [&](const W Eq) {
        W D0 = (((Eq & VP) + VP) ^ VP) | Eq | VN;
        if (TRANS) {
            D0 |= ~D0prev & (Eq << 1) & (Eqprev >> 1);
            D0prev = D0;
            Eqprev = Eq;
        }
        const W HP = VN | ~(D0 | VP);
        const W HN = D0 & VP;
        const W X = D0 >> 1;
        VN = X & HP;
        VP = HN | ~(X | HP);
        acc += D0 & emask;
    };
    // fetch the next 16 pattern-stream bytes into the top of the window
    auto take_pattern = [&]() {
        uint32_t x[4];
        sa.take(x);
        if (PLANES) {
            g7prev = g7cur;
            g7cur = gather_bits16(x, 7);
        }
#pragma unroll
        for (int w = 0; w < 4; w++) win[NW - 4 + w] = x[w] & CMASK;
    };
    auto slide_window = [&]() {
#pragma unroll
        for (int w = 0; w + 4 < NW; w++) win[w] = win[w + 4];
    };

    // 16 columns whose first column sits at circular position PH (0 or 16).  The read-modify-write that ORs byte
    // j + L in is software-pipelined by hand: its load is issued one column early (right after the previous store,
    // which it must follow because two entering bytes may share an entry), so neither the store nor the look-up that
    // follows it waits for shared-memory latency.
    auto superstep = [&](auto phc) {
        constexpr uint32_t PH = (uint32_t)decltype(phc)::value;
        take_pattern();
        sb.take(bw);
#pragma unroll
        for (int w = 0; w < 4; w++) bc[w] = bw[w] & CMASK;
        // start of chunk q: planes of the entering bytes (their positions' old bits are dead), dead block zeroed by value
        auto begin_chunk = [&](const int q) {
            const uint32_t pos = (PH + (uint32_t)(C * q + L)) & 31u;  // circular position of the dying / entering block
            if (PLANES) {
                const W blk = (C == 16 ? 0xffffu : 0xffu) << pos;
                W tops;
                if (C == 16)
                    tops = g7cur << pos;
                else
                    tops = (q == 0 ? (g7prev >> 8) : (g7cur & 0xffu)) << pos;
                A7 = (A7 & ~blk) | (tops & blk);
            }
#pragma unroll
            for (int t = 0; t < C; t++) {
                const int o = 16 - C + C * q + t;  // byte offset of the dead byte inside the window
                uint8_t *p = blk_entry(tab, win[o >> 2], o & 3, pitch) + pos / 8;
                if (C == 16)
                    *(uint16_t *)p = 0;
                else
                    *p = 0;
            }
        };
        auto enter_addr = [&](const int u) {  // entry of stream byte j + L for column u of this superstep
            const int o = 16 + u + L;
            return (W *)blk_entry(tab, win[o >> 2], o & 3, pitch);
        };
        begin_chunk(0);
        W *pend = enter_addr(0);
        W pend_val = *pend;
#pragma unroll
        for (int u = 0; u < 16; u++) {
            *pend = pend_val | (1u << ((PH + (uint32_t)(u + L)) & 31u));
            W raw = *(const W *)blk_entry(tab, bc[u >> 2], u & 3, pitch);
            if (PLANES) raw &= ~(A7 ^ prmt(bw[u >> 2], 0u, 0x8888u | (uint32_t)((u & 3) * 0x1111)));
            if (u + 1 < 16) {
                if ((u + 1) % C == 0) begin_chunk((u + 1) / C);
                pend = enter_addr(u + 1);
                pend_val = *pend;
            }
            step(funnel_r(raw, raw, (PH + (uint32_t)u) & 31u));
        }
        matches += acc >> e;
        acc = 0;
        slide_window();
    };

    // Early exit: the value on the final diagonal, diff + j - matches(j), never decreases from column to column (its
    // delta is the D0 bit: 0 or +1) and ends as the result, so once it exceeds max_k the answer is "> max_k" whatever
    // follows.  Tested once per 32 columns, at the loop's own back edge, against the count as of the START of the pass:
    // that value is long computed, so the branch never waits for the recurrence's dependency chain (testing the fresh
    // count drained the pipeline at every back edge, a test between the two unrolled supersteps split the
    // straight-line body: 3-9 % on the matching pairs).  Unrelated pairs leave after two passes.
    // The exit is folded into the loop bound (lim drops to 0) so that the loop keeps a single exit.
    int lim = n;
    int j0 = 0;
    for (; j0 + 32 <= lim; j0 += 32) {
        const uint32_t m_old = matches;
        superstep(IntC<0>());
        superstep(IntC<16>());
        lim = (uint32_t)diff + (uint32_t)j0 - m_old > max_k ? 0 : lim;
    }
    const bool dead = lim == 0;  // (n >= 1 here)
    prefetch_next(next);
    uint32_t phase = 0;
    if (!dead && j0 + 16 <= n) {
        superstep(IntC<0>());
        j0 += 16;
        phase = 16;
    }
    // the table now holds stream bytes [j0 - C, j0 + L) = window bytes from offset 16 - C
    constexpr int SW0 = (16 - C) / 4;
    uint32_t sw[NW - SW0];
    if (!dead && j0 < n) {  // last n % 16 columns: slide the table column by column (exact clear + set), rolled
        take_pattern();
        sb.take(bw);
#pragma unroll
        for (int w = 0; w < NW - SW0; w++) sw[w] = win[SW0 + w];
        W tops = 0;
        if (PLANES) {
            if (C == 16)
                tops = g7cur << ((phase + 16u) & 31u);
            else
                tops = ((g7prev >> 8) << ((phase + 24u) & 31u)) | ((g7cur & 0xffu) << phase);
        }
        for (int u = 0; u < n - j0; u++) {
            // byte j - C (sw byte 0) is dead and byte j + L (sw byte 32) takes over its position
            const W bit = 1u << ((phase + (uint32_t)(u + L)) & 31u);
            *(W *)blk_entry(tab, sw[0] & 0xffu, 0, pitch) &= ~bit;
            *(W *)blk_entry(tab, sw[8] & 0xffu, 0, pitch) |= bit;
            if (PLANES) A7 = (A7 & ~bit) | (tops & bit);
            const uint32_t bch = bw[0] & 0xffu;
            W raw = *(const W *)blk_entry(tab, bch & (CMASK & 0xffu), 0, pitch);
            if (PLANES) raw &= ~(A7 ^ (0u - (bch >> 7)));
            step(funnel_r(raw, raw, (phase + (uint32_t)u) & 31u));
#pragma unroll
            for (int w = 0; w < 3; w++) bw[w] = funnel_r(bw[w], bw[w + 1], 8);
            bw[3] >>= 8;
#pragma unroll
            for (int w = 0; w + 1 < NW - SW0; w++) sw[w] = funnel_r(sw[w], sw[w + 1], 8);
            sw[NW - 1 - SW0] >>= 8;
        }
        matches += acc >> e;
    } else {
#pragma unroll
        for (int w = 0; w < NW - SW0; w++) sw[w] = win[SW0 + w];
    }
    // leave the table clean: it holds exactly the 32 stream bytes sw[0..7]
#pragma unroll
    for (int t = 0; t < 32; t++) *(W *)blk_entry(tab, sw[t >> 2], t & 3, pitch) = 0;
    return dead ? 0xFFFFFFFFu : (uint32_t)diff + (uint32_t)n - matches;
}

template <bool TRANS, int PLANES, int C>
TA_HD uint32_t pair_unit_costs_blk(const uint8_t *a, uint64_t a_len, const uint8_t *b, uint64_t b_len, uint32_t k,
                                   uint8_t *tab, const uint32_t pitch, const NextHint *next = nullptr) {
    if (a_len > b_len) {
        const uint8_t *tp = a;
        a = b;
        b = tp;
        const uint64_t tl = a_len;
        a_len = b_len;
        b_len = tl;
    }
    const int m = (int)a_len, n = (int)b_len;
    const uint32_t diff = (uint32_t)(n - m);
    const uint32_t max_k = k < (uint32_t)n ? k : (uint32_t)n;
    if (diff > max_k) return 0xFFFFFFFFu;
    if (m == 0) return (uint32_t)n;
    const uint32_t d = distance_blk<TRANS, PLANES, C>(a, m, b, n, max_k, tab, pitch, next);
    return d <= max_k ? d : 0xFFFFFFFFu;
}

// ---------------------------------------------------------------------------------------------------------------
// distance_duo: TWO pairs per thread for bands of W <= 9 diagonals (unit costs without transpositions, max_k <= 8).
//
// The block-table scheme of distance_blk with a 16-position frame (C = 8, L = 8) needs only 16 bits per pair, so two
// pairs share every 32-bit word: pair P owns bits [16 P, 16 P + 16) of the table entries (each pair reads and writes
// only its own half-word, so the two never interfere), of the A7 plane and of VP / VN / D0.  The recurrence, the
// plane test, the rotation and the match counting are then issued ONCE for both pairs; only the byte extraction and
// the table accesses stay per pair.  The packed recurrence is two exact, independent 16-bit recurrences:
//   * bit 15 of each half is masked out of the addends ((Eq & VP & M) + (VP & M), M = 0x7fff7fff), so no carry leaves
//     a half; bit 15 of the sum then holds the incoming carry c14 and D0 = ((S ^ VPm) | Eq | VN) still equals the
//     16-bit result at that bit (if Eq15 = 0 the 16-bit word has (A + VP) ^ VP = c14 there as well; if Eq15 = 1, D0 = 1);
//   * X = (D0 >> 1) & M shifts a zero into row 15 of each half, as a 16-bit shift would.
// Both pairs must have the same number of full 16-column supersteps (n_A / 16 == n_B / 16); the last columns run in a
// rolled loop that counts each pair's matches only inside its own length.  dP = exact distance if <= max_k.
struct DuoSide {  // per-pair state of the duo kernel
    Stream sa, sb;
    uint32_t win[12];  // class bits of pattern-stream bytes [16 s - 16, 16 s + 32)
    uint32_t bw[4], bc[4];
    uint32_t g7prev, g7cur;
    int n, diff, e, dhi;
};

TA_HD void distance_duo(const uint8_t *aA, int mA, const uint8_t *bA, int nA, uint32_t maxkA, const uint8_t *aB, int mB,
                        const uint8_t *bB, int nB, uint32_t maxkB, uint8_t *tab, const uint32_t pitch, uint32_t &dA,
                        uint32_t &dB, const NextHint *next = nullptr) {
    constexpr uint32_t CMASK = 0x7f7f7f7fu, M = 0x7fff7fffu;
    DuoSide sd[2];
    auto setup = [&](DuoSide &x, const uint8_t *a, int m, const uint8_t *b, int n, uint32_t max_k) {
        x.n = n;
        x.diff = n - m;
        x.e = (int)((max_k - (uint32_t)x.diff) >> 1);
        x.dhi = x.diff + x.e;
        x.sa.init((intptr_t)a - x.dhi, (uintptr_t)a, (uintptr_t)a + m - 1);
        x.sb.init((intptr_t)b, (uintptr_t)b, (uintptr_t)b + n - 1);
#pragma unroll
        for (int w = 0; w < 12; w++) x.win[w] = 0;
        x.g7prev = x.g7cur = 0;
    };
    setup(sd[0], aA, mA, bA, nA, maxkA);
    setup(sd[1], aB, mB, bB, nB, maxkB);
    auto half_ptr = [&](int P, uint32_t w, int t) { return (uint16_t *)(tab + byte_of(w, t) * pitch + 2 * P); };

    uint32_t A7 = 0, VP = 0, emask = 0;
#pragma unroll
    for (int P = 0; P < 2; P++) {
        DuoSide &x = sd[P];
        uint32_t t0[4];
        x.sa.take(t0);  // T_0
        x.g7cur = gather_bits16(t0, 7);
        A7 |= (x.g7cur & 0xffu) << (16 * P);
#pragma unroll
        for (int w = 0; w < 4; w++) x.win[4 + w] = t0[w] & CMASK;
#pragma unroll
        for (int t = 0; t < 8; t++) *half_ptr(P, x.win[4 + (t >> 2)], t & 3) |= (uint16_t)(1u << t);
        const uint32_t vp16 = x.dhi >= 16 ? 0u : ((0xffffu << x.dhi) & 0xffffu);
        VP |= vp16 << (16 * P);
        emask |= (1u << x.e) << (16 * P);
    }
    uint32_t VN = ~VP;
    uint32_t acc = 0, matches[2] = {0, 0};

    auto step = [&](const uint32_t Eq, const uint32_t cmask) {
        const uint32_t VPm = VP & M;
        const uint32_t S = (Eq & VPm) + VPm;
        const uint32_t D0 = (S ^ VPm) | Eq | VN;
        const uint32_t HP = VN | ~(D0 | VP);
        const uint32_t Ds = D0 >> 1;
        const uint32_t Y = ~((Ds & M) | HP);
        VN = Ds & M & HP;
        VP = (D0 & VP) | Y;
        acc += D0 & cmask;
    };
    auto take_side = [&](DuoSide &x) {
        uint32_t t[4];
        x.sa.take(t);
        x.g7prev = x.g7cur;
        x.g7cur = gather_bits16(t, 7);
#pragma unroll
        for (int w = 0; w < 4; w++) x.win[8 + w] = t[w] & CMASK;
        x.sb.take(x.bw);
#pragma unroll
        for (int w = 0; w < 4; w++) x.bc[w] = x.bw[w] & CMASK;
    };

    const int steps = nA >> 4;  // == nB >> 4
    bool dead[2] = {false, false};
    int s_done = steps;
    int slim = steps;
    for (int s = 0; s < slim; s++) {
        const uint32_t m_old0 = matches[0], m_old1 = matches[1];
        take_side(sd[0]);
        take_side(sd[1]);
        // start of chunk q (8 columns): planes of the entering block, dead block [g - 8, g) zeroed by value
        auto begin_chunk = [&](const int q) {
#pragma unroll
            for (int P = 0; P < 2; P++) {
                DuoSide &x = sd[P];
                const uint32_t blk = (q == 0 ? 0xff00u : 0x00ffu) << (16 * P);
                const uint32_t tops = (q == 0 ? x.g7prev : x.g7cur) << (16 * P);
                A7 = (A7 & ~blk) | (tops & blk);
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const int o = 8 + 8 * q + t;  // window offset of the dead byte; its block sits in byte (1 - q) of the half
                    *((uint8_t *)half_ptr(P, x.win[o >> 2], o & 3) + (1 - q)) = 0;
                }
            }
        };
        auto enter_ptr = [&](const int P, const int u) {  // half-entry of stream byte j + 8 for column u
            const int o = 24 + u;
            return half_ptr(P, sd[P].win[o >> 2], o & 3);
        };
        begin_chunk(0);
        uint16_t *pend[2] = {enter_ptr(0, 0), enter_ptr(1, 0)};
        uint32_t pend_val[2] = {*pend[0], *pend[1]};
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const uint32_t bit = 1u << ((u + 8) & 15);
            *pend[0] = (uint16_t)(pend_val[0] | bit);
            *pend[1] = (uint16_t)(pend_val[1] | bit);
            const uint32_t rawA = *half_ptr(0, sd[0].bc[u >> 2], u & 3);
            const uint32_t rawB = *half_ptr(1, sd[1].bc[u >> 2], u & 3);
            uint32_t raw = prmt(rawA, rawB, 0x5410u);
            const uint32_t i = (uint32_t)(u & 3);
            raw &= ~(A7 ^ prmt(sd[0].bw[u >> 2], sd[1].bw[u >> 2], (0x8u | i) | ((0x8u | i) << 4) | ((0xcu | i) << 8) | ((0xcu | i) << 12)));
            if (u + 1 < 16) {
                if (u + 1 == 8) begin_chunk(1);
                pend[0] = enter_ptr(0, u + 1);
                pend[1] = enter_ptr(1, u + 1);
                pend_val[0] = *pend[0];
                pend_val[1] = *pend[1];
            }
            uint32_t Eq = raw;
            if (u != 0) {
                const uint32_t m1 = (0xffffu >> u) * 0x10001u;
                Eq = ((raw >> u) & m1) | ((raw << (16 - u)) & ~m1);
            }
            step(Eq, emask);
        }
        matches[0] += (acc & 0xffffu) >> sd[0].e;
        matches[1] += (acc >> 16) >> sd[1].e;
        acc = 0;
#pragma unroll
        for (int P = 0; P < 2; P++)
#pragma unroll
            for (int w = 0; w < 8; w++) sd[P].win[w] = sd[P].win[w + 4];
        // early exit (see distance_blk): a pair whose final-diagonal value passed max_k is decided; leave when both
        // are.  The test uses the counts as of the start of this superstep, which are long computed.
        const uint32_t cols = (uint32_t)s << 4;
        dead[0] = dead[0] || (uint32_t)sd[0].diff + cols - m_old0 > maxkA;
        dead[1] = dead[1] || (uint32_t)sd[1].diff + cols - m_old1 > maxkB;
        if (dead[0] && dead[1]) {  // folded into the loop bound: the loop keeps a single exit
            s_done = s + 1;
            slim = 0;
        }
    }
    prefetch_next(next);
    // tables now hold stream bytes [j0 - 8, j0 + 8) of each pair = window bytes from offset 8
    const int j0 = s_done << 4;
    const bool both_dead = dead[0] && dead[1];
    const int rA = nA - j0, rB = nB - j0, r = both_dead ? 0 : (rA > rB ? rA : rB);
    uint32_t sw[2][10];
    if (r > 0) {  // last columns: exact per-column sliding, rolled; each pair counts only its own columns
        take_side(sd[0]);
        take_side(sd[1]);
        uint32_t tops = 0;
#pragma unroll
        for (int P = 0; P < 2; P++) {
            tops |= ((sd[P].g7prev & 0xff00u) | (sd[P].g7cur & 0x00ffu)) << (16 * P);
#pragma unroll
            for (int w = 0; w < 10; w++) sw[P][w] = sd[P].win[2 + w];
        }
        for (int u = 0; u < r; u++) {
            const uint32_t bit = 1u << ((u + 8) & 15);
#pragma unroll
            for (int P = 0; P < 2; P++) {
                // byte j - 8 (sw byte 0) is dead and byte j + 8 (sw byte 16) takes over its position
                *half_ptr(P, sw[P][0] & 0xffu, 0) &= (uint16_t)~bit;
                *half_ptr(P, sw[P][4] & 0xffu, 0) |= (uint16_t)bit;
            }
            const uint32_t bit2 = bit * 0x10001u;
            A7 = (A7 & ~bit2) | (tops & bit2);
            const uint32_t chA = sd[0].bw[0] & 0xffu, chB = sd[1].bw[0] & 0xffu;
            uint32_t raw = (uint32_t)*half_ptr(0, chA & 0x7fu, 0) | ((uint32_t)*half_ptr(1, chB & 0x7fu, 0) << 16);
            raw &= ~(A7 ^ (((0u - (chA >> 7)) & 0xffffu) | ((0u - (chB >> 7)) << 16)));
            const uint32_t m1 = (0xffffu >> u) * 0x10001u;
            const uint32_t Eq = u ? (((raw >> u) & m1) | ((raw << (16 - u)) & ~m1)) : raw;
            step(Eq, (u < rA ? (emask & 0xffffu) : 0u) | (u < rB ? (emask & 0xffff0000u) : 0u));
#pragma unroll
            for (int P = 0; P < 2; P++) {
#pragma unroll
                for (int w = 0; w < 3; w++) sd[P].bw[w] = funnel_r(sd[P].bw[w], sd[P].bw[w + 1], 8);
                sd[P].bw[3] >>= 8;
#pragma unroll
                for (int w = 0; w < 9; w++) sw[P][w] = funnel_r(sw[P][w], sw[P][w + 1], 8);
                sw[P][9] >>= 8;
            }
        }
        matches[0] += (acc & 0xffffu) >> sd[0].e;
        matches[1] += (acc >> 16) >> sd[1].e;
    } else {
#pragma unroll
        for (int P = 0; P < 2; P++)
#pragma unroll
            for (int w = 0; w < 10; w++) sw[P][w] = sd[P].win[2 + w];
    }
    // leave the table clean: each half holds exactly the 16 stream bytes sw[P][0..3]
#pragma unroll
    for (int P = 0; P < 2; P++)
#pragma unroll
        for (int t = 0; t < 16; t++) *half_ptr(P, sw[P][t >> 2], t & 3) = 0;
    dA = dead[0] ? 0xFFFFFFFFu : (uint32_t)sd[0].diff + (uint32_t)nA - matches[0];
    dB = dead[1] ? 0xFFFFFFFFu : (uint32_t)sd[1].diff + (uint32_t)nB - matches[1];
}

// One pair's contract up to the point where the DP is needed (reference src/levenshtein.rs:386-430 with unit
// costs): swaps so that a is the shorter string, clamps k; returns false when the answer is already known (*out).
TA_HD bool unit_costs_prepare(const uint8_t *&a, uint64_t &a_len, const uint8_t *&b, uint64_t &b_len, uint32_t k,
                              uint32_t &max_k, uint32_t *out) {
    if (a_len > b_len) {
        const uint8_t *tp = a;
        a = b;
        b = tp;
        const uint64_t tl = a_len;
        a_len = b_len;
        b_len = tl;
    }
    const uint32_t diff = (uint32_t)(b_len - a_len);
    max_k = k < (uint32_t)b_len ? k : (uint32_t)b_len;
    if (diff > max_k) {
        *out = 0xFFFFFFFFu;
        return false;
    }
    if (a_len == 0) {
        *out = (uint32_t)b_len;
        return false;
    }
    return true;
}

template <bool TRANS, int PLANES, typename W>
TA_HD uint32_t pair_unit_costs_tab(const uint8_t *a, uint64_t a_len, const uint8_t *b, uint64_t b_len, uint32_t k,
                                   uint8_t *tab, const uint32_t pitch) {
    if (a_len > b_len) {
        const uint8_t *tp = a;
        a = b;
        b = tp;
        const uint64_t tl = a_len;
        a_len = b_len;
        b_len = tl;
    }
    const int m = (int)a_len, n = (int)b_len;
    const uint32_t diff = (uint32_t)(n - m);
    const uint32_t max_k = k < (uint32_t)n ? k : (uint32_t)n;
    if (diff > max_k) return 0xFFFFFFFFu;
    if (m == 0) return (uint32_t)n;
    const uint32_t d = distance_tab<TRANS, PLANES, W>(a, m, b, n, max_k, tab, pitch);
    return d <= max_k ? d : 0xFFFFFFFFu;
}

// The whole per-pair contract of levenshtein_naive_k_with_opts for unit costs (reference
// src/levenshtein.rs:386-430, 539-541): swap so a is shorter, clamp k, early None, then the banded distance.
// can_handle(k, trans, max_len) on the host guarantees the band fits 32 rows.
template <bool TRANS>
TA_HD uint32_t pair_unit_costs(const uint8_t *a, uint64_t a_len, const uint8_t *b, uint64_t b_len, uint32_t k) {
    if (a_len > b_len) {
        const uint8_t *tp = a;
        a = b;
        b = tp;
        const uint64_t tl = a_len;
        a_len = b_len;
        b_len = tl;
    }
    const int m = (int)a_len, n = (int)b_len;
    const uint32_t diff = (uint32_t)(n - m);
    // max_k = min(k, min(m*1, 2m*1) + diff*1) = min(k, n)   (src/levenshtein.rs:400-423 with unit costs)
    const uint32_t max_k = k < (uint32_t)n ? k : (uint32_t)n;
    if (diff > max_k) return 0xFFFFFFFFu;  // unit_k = max_k (src/levenshtein.rs:426-430)
    if (m == 0) return (uint32_t)n;        // n <= max_k here
    const uint32_t d = distance32<TRANS>(a, m, b, n, max_k);
    return d <= max_k ? d : 0xFFFFFFFFu;
}

}  // namespace bitpar
