// lev_fr.cu -- unit-cost k-bounded distance by diagonal extension ("furthest reaching"), for long strings on sm_100a.
//
// Contract: levenshtein_simd_k_with_opts / levenshtein_naive_k_with_opts (reference src/levenshtein.rs:714-827,
// 376-545) for LEVENSHTEIN_COSTS and RDAMERAU_COSTS: out = d if d <= k else TA_NONE.  Same answers as the
// bit-parallel kernels in lev_bitpar.cu; the dispatcher (ta_launch_lev) picks this kernel when the strings are long
// compared with k^2, where filling the band costs 30 instructions per column and this costs ~0.4 per byte.
//
// Algorithm (Ukkonen / Landau-Vishkin / Myers O(ND); executable model: tests/test_fr_model.py).  With unit costs the
// values on a diagonal c = j - i of the DP matrix never decrease, so level e is described by FR_e[c] = the furthest
// row whose cell on diagonal c costs <= e.  FR_e[c] = slide(max(FR_{e-1}[c] + 1, FR_{e-1}[c-1], FR_{e-1}[c+1] + 1,
// FR_{e-1}[c] + 2 if the two bytes after it are swapped in b)), clamped to the matrix, where slide() runs down the
// diagonal while a[row] == b[row + c].  The answer is the first e with FR_e[|b|-|a|] = |a|; levels stop at max_k.
// Only diagonals that can still reach the target within max_k are kept: |c - diff| <= max_k - e, |c| <= e.
//
// Mapping: 8 lanes (an "octet") per pair, 4 pairs per warp, persistent octets striding over the batch.
//  * a ROUND gives each lane one diagonal of the current level: neighbours from shared memory (two ping-pong
//    arrays of 2 max_k + 3 slots per octet), then the 16-byte chunk of `a` that holds the start row is compared with
//    the matching 16 bytes of `b` (one aligned LDG.128 against three aligned LDG.64 re-aligned with funnel shifts);
//    most diagonals stop inside that chunk.
//  * a diagonal that matched to the end of its chunk (the alignment's own path, as a rule) is extended by whole
//    octets: lane l takes chunk l of the next 128 bytes -- one full line of each string per step, coalesced -- and
//    the lane with the smallest stopping row (one redux.sync) ends the slide.  Octets of the warp that have no slide
//    of their own serve the others' (block j of every `power` blocks: a warp with one long slide moves 512 bytes
//    of each string per step); measured lane utilisation of the slide loop went from 11 to ~27 of 32 lanes.
//  Algorithmic bytes per pair: |a| + |b| + 4 (every byte on the alignment's path is read once).
//  Tried and dropped (round 2, profiles/README.md): a lane per pair with one-byte probes and the octet serving its eight
//  lanes' slides -- rounds become ~20 instructions per diagonal of one lane, but a lane's probes are a serial chain of
//  scattered global loads (one L2 round trip per diagonal) and the lanes of a warp are rarely in the same phase
//  (4 of 32 lanes active in the probe loop): 1.74 ms / 1.94 ms with queued slides against 1.10 ms for this form.
#include <stdlib.h>

#include <algorithm>

#include "ta_common.cuh"

namespace {

constexpr int FR_NEG = -(1 << 30);
constexpr unsigned FULL = 0xFFFFFFFFu;

// Bytes [r, r + 16) of the 24-byte window v0:v1:v2 (three aligned 8-byte loads, r < 8) XOR the 16 bytes of A.
__device__ __forceinline__ void xor16(const uint4 A, const uint2 v0, const uint2 v1, const uint2 v2, const uint32_t r,
                                      uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3) {
    const uint32_t bs = (r & 3u) * 8u;
    const uint32_t g0 = __funnelshift_r(v0.x, v0.y, bs), g1 = __funnelshift_r(v0.y, v1.x, bs),
                   g2 = __funnelshift_r(v1.x, v1.y, bs), g3 = __funnelshift_r(v1.y, v2.x, bs),
                   g4 = __funnelshift_r(v2.x, v2.y, bs);
    const bool w1 = (r & 4u) != 0;
    x0 = A.x ^ (w1 ? g1 : g0), x1 = A.y ^ (w1 ? g2 : g1), x2 = A.z ^ (w1 ? g3 : g2), x3 = A.w ^ (w1 ? g4 : g3);
}

// index of the first non-zero byte of x0:x1:x2:x3 (16 when there is none)
__device__ __forceinline__ uint32_t first_nonzero_byte(const uint32_t x0, const uint32_t x1, const uint32_t x2,
                                                       const uint32_t x3) {
    const bool lo_any = (x0 | x1) != 0u;
    const uint32_t lo = x0 ? x0 : x1, hi = x2 ? x2 : x3;
    const uint32_t lo_base = x0 ? 0u : 4u, hi_base = x2 ? 8u : 12u;
    const uint32_t w = lo_any ? lo : hi, base = lo_any ? lo_base : hi_base;
    return w ? base + ((uint32_t)(__ffs((int)w) - 1) >> 3) : 16u;
}

// the three aligned 8-byte words that hold bytes [pbs, pbs + 16) of b's 8-byte frame; words outside [0, last8] are not
// loaded (they can only supply bytes outside the string, which the caller never looks at)
__device__ __forceinline__ void load_b24(const uint2 *__restrict__ b8, const int pbs, const int last8, uint2 &v0,
                                         uint2 &v1, uint2 &v2) {
    const int j = pbs >> 3;
    v0 = v1 = v2 = make_uint2(0u, 0u);
    if (j >= 0 && j <= last8) v0 = __ldg(b8 + j);
    if (j + 1 >= 0 && j + 1 <= last8) v1 = __ldg(b8 + j + 1);
    if (j + 2 >= 0 && j + 2 <= last8) v2 = __ldg(b8 + j + 2);
}

__device__ __forceinline__ void prefetch_l2_line(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// What an octet publishes for the octets that help it slide (shared memory, one slot per octet)
struct __align__(16) FrSlot {
    const uint4 *a16;
    const uint2 *b8;
    int a_mis, b_mis, last8, pad0;
    int sc, srow, slim, pad1;
};

// 8 CTAs per SM = 64 registers, no spills (measured: 1.065 ms at 8, 1.098 at 9, 1.198 at 10, 1.288 at 12 CTAs per SM on
// 256 Ki pairs of 4096 bytes: more resident warps do not pay for the spills they cost)
template <bool TRANS>
__global__ void __launch_bounds__(128, 8) lev_fr_kernel(const uint8_t *__restrict__ a, const uint64_t *__restrict__ a_off,
                                                     const uint8_t *__restrict__ b, const uint64_t *__restrict__ b_off,
                                                     const uint32_t *__restrict__ idx, size_t n, uint32_t k,
                                                     uint32_t slots_pad, uint32_t *__restrict__ out) {
    extern __shared__ __align__(16) int fr_all[];  // [octet][2][slots_pad], then one FrSlot per octet
    // Who serves which slide, for every set `amask` of octets (of a warp) that own one and every octet `ow`: the i-th
    // idle octet helps the (i mod nact)-th active one with block 1 + i / nact of each step.
    __shared__ uint32_t lut[64];
    if (threadIdx.x < 64) {
        const uint32_t am = threadIdx.x >> 2, o_me = threadIdx.x & 3u;
        uint32_t info = 0;
        if (am) {
            const int nact = __popc(am), nidle = 4 - nact;
            int tgt_i, boff;
            if ((am >> o_me) & 1u) {
                tgt_i = __popc(am & ((1u << o_me) - 1u)), boff = 0;
            } else {
                const int rank = __popc(~am & ((1u << o_me) - 1u) & 0xFu);
                tgt_i = rank % nact, boff = 1 + rank / nact;
            }
            const int power = 1 + nidle / nact + (tgt_i < nidle % nact ? 1 : 0);
            const int tgt = (int)__fns(am, 0, tgt_i + 1);  // warp-local index of the octet that owns the slide
            uint32_t gm = 0;
            for (int o = 0; o < 4; o++) {
                const int ti = ((am >> o) & 1u) ? __popc(am & ((1u << o) - 1u)) : __popc(~am & ((1u << o) - 1u) & 0xFu) % nact;
                if (ti == tgt_i) gm |= 1u << o;
            }
            info = (uint32_t)tgt | ((uint32_t)boff << 2) | ((uint32_t)power << 4) | (gm << 8);
        }
        lut[threadIdx.x] = info;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, l = lane & 7u, osh = lane & 24u, ow = lane >> 3;
    const uint32_t oct = threadIdx.x >> 3, noct = blockDim.x >> 3;
    int *const fr = fr_all + (size_t)oct * 2u * slots_pad;
    FrSlot *const slots = (FrSlot *)(fr_all + (size_t)noct * 2u * slots_pad);
    FrSlot *const wslots = slots + (oct & ~3u);  // the four slots of this warp
    const size_t stride = (size_t)gridDim.x * noct;
    size_t w = (size_t)blockIdx.x * noct + oct;

    // octet-uniform state (every lane of the octet holds the same values)
    bool have = false, drained = false;
    const uint8_t *pa = nullptr, *pb = nullptr;
    const uint4 *a16 = nullptr;
    const uint2 *b8 = nullptr;
    uint32_t a_mis = 0, b_mis = 0;
    int last8 = 0;
    int m = 0, nn = 0, diff = 0, max_k = 0, e = 0, hi = 0, cbase = 0;
    size_t pair = 0;

    for (;;) {
        // ---- A: an idle octet takes its next pair (answers known without DP are written at once) -----------------
        while (!have && !drained) {
            if (w >= n) {
                drained = true;
                break;
            }
            pair = idx ? (size_t)idx[w] : w;
            w += stride;
            uint64_t a0 = a_off[pair], b0 = b_off[pair];
            uint64_t la = a_off[pair + 1] - a0, lb = b_off[pair + 1] - b0;
            pa = a + a0, pb = b + b0;
            if (la > lb) {  // a is the shorter string (reference src/levenshtein.rs:386-390)
                const uint8_t *tp = pa;
                pa = pb, pb = tp;
                const uint64_t tl = la;
                la = lb, lb = tl;
            }
            m = (int)la, nn = (int)lb, diff = nn - m;
            max_k = (int)(k < (uint32_t)nn ? k : (uint32_t)nn);
            if (diff > max_k) {
                if (l == 0) out[pair] = 0xFFFFFFFFu;
                continue;
            }
            if (m == 0) {
                if (l == 0) out[pair] = (uint32_t)nn;
                continue;
            }
            a_mis = (uint32_t)((uintptr_t)pa & 15u), b_mis = (uint32_t)((uintptr_t)pb & 7u);
            a16 = (const uint4 *)(pa - a_mis), b8 = (const uint2 *)(pb - b_mis);
            last8 = (int)((b_mis + (uint32_t)nn - 1u) >> 3);
            const int4 neg = make_int4(FR_NEG, FR_NEG, FR_NEG, FR_NEG);
            for (uint32_t q = l * 4u; q < 2u * slots_pad; q += 32u) *(int4 *)(fr + q) = neg;
            if (l == 0) {
                FrSlot &sl = slots[oct];
                sl.a16 = a16, sl.b8 = b8, sl.a_mis = (int)a_mis, sl.b_mis = (int)b_mis, sl.last8 = last8;
            }
            e = 0, hi = 0, cbase = 0;
            have = true;
        }
        if (__all_sync(FULL, !have)) break;
        __syncwarp();

        // ---- B: one round = the next 8 diagonals of the current level, one per lane --------------------------------
        const int c = cbase + (int)l;
        const bool valid = have && c <= hi;
        const int s = c + max_k + 1;
        const int lim = min(m, nn - c);
        int *const cur = fr + (e & 1) * (int)slots_pad;
        int v = 0, row_next = 0;
        bool pend = false;
        if (valid) {
            int t = 0;
            if (e > 0) {
                const int *const prev = fr + ((e & 1) ^ 1) * (int)slots_pad;
                const int L = prev[s - 1], M = prev[s], R = prev[s + 1];
                t = max(max(M + 1, L), R + 1);
                if (TRANS && M >= 0 && M + 2 <= lim) {  // restricted transposition out of the furthest cell
                    if (pa[M] == pb[M + c + 1] && pa[M + 1] == pb[M + c]) t = max(t, M + 2);
                }
            }
            t = min(t, lim);
            v = t;
            if (t < lim) {
                // the 16-byte chunk of `a` that holds row t, against the bytes of `b` on this diagonal
                const uint32_t pos = a_mis + (uint32_t)t, skip = pos & 15u;
                const int row0 = t - (int)skip;
                const int pbs = (int)b_mis + row0 + c;  // byte position in b's aligned frame (< 0: skipped bytes only)
                const uint4 A = __ldg(a16 + (pos >> 4));
                uint2 v0, v1, v2;
                load_b24(b8, pbs, last8, v0, v1, v2);
                uint32_t x0, x1, x2, x3;
                xor16(A, v0, v1, v2, (uint32_t)pbs & 7u, x0, x1, x2, x3);
                const uint32_t s8 = skip * 8u;  // ignore the bytes below `skip`
                x0 = s8 >= 32u ? 0u : x0 & (0xFFFFFFFFu << s8);
                x1 = s8 >= 64u ? 0u : (s8 > 32u ? x1 & (0xFFFFFFFFu << (s8 - 32u)) : x1);
                x2 = s8 >= 96u ? 0u : (s8 > 64u ? x2 & (0xFFFFFFFFu << (s8 - 64u)) : x2);
                x3 = s8 > 96u ? x3 & (0xFFFFFFFFu << (s8 - 96u)) : x3;
                const uint32_t off = first_nonzero_byte(x0, x1, x2, x3);
                v = min(row0 + (int)off, lim);
                row_next = row0 + 16;
                pend = off == 16u && row_next < lim;
            }
        }
        uint32_t P = (__ballot_sync(FULL, pend) >> osh) & 0xFFu;

        // ---- C: diagonals that matched to the end of their chunk are extended 128 bytes per octet and step.  Octets
        //      without a slide of their own help the others (block j of every `power` blocks), so a warp with one
        //      long slide moves 512 bytes of each string per step.
        if (__any_sync(FULL, P != 0u)) {
            bool in_slide = false, serving = false;
            bool assign = true;  // warp-uniform: (re)assign octets to slides before the next step
            uint32_t q = 0;
            int own_c = 0, own_row = 0, own_lim = 0;
            // the slide this octet serves (its own or another octet's)
            const uint4 *h_a16 = a16;
            const uint2 *h_b8 = b8;
            int h_amis = 0, h_bmis = 0, h_last8 = 0, h_c = 0, h_row = 0, h_lim = 0, h_step = 128;
            uint32_t gmask = 0xFFu << osh, g_boff = 0;
            for (;;) {
                if (assign) {
                    assign = false;
                    // octets without a slide take their next pending diagonal (lane qq's start row, diagonal, limit)
                    const bool start = !in_slide && P != 0u;
                    if (__any_sync(FULL, start)) {
                        const uint32_t qq = P ? (uint32_t)__ffs((int)P) - 1u : 0u;
                        const int c2 = __shfl_sync(FULL, c, (int)qq, 8), r2 = __shfl_sync(FULL, row_next, (int)qq, 8),
                                  l2 = __shfl_sync(FULL, lim, (int)qq, 8);
                        if (start) q = qq, own_c = c2, own_row = r2, own_lim = l2, in_slide = true;
                    }
                    const uint32_t bal = __ballot_sync(FULL, in_slide);
                    const uint32_t amask = ((bal & 0x01010101u) * 0x01020408u) >> 24;  // bit o: octet o owns a slide
                    if (amask == 0u) break;
                    if (in_slide && l == 0) {
                        FrSlot &sl = slots[oct];
                        sl.sc = own_c, sl.srow = own_row, sl.slim = own_lim;
                    }
                    __syncwarp();
                    const uint32_t info = lut[amask * 4u + ow];  // tgt | boff << 2 | power << 4 | group octets << 8
                    const int tgt = (int)(info & 3u), boff = (int)((info >> 2) & 3u), power = (int)((info >> 4) & 7u);
                    gmask = (((info >> 8) & 0xFu) * 0x00204081u & 0x01010101u) * 0xFFu;
                    const FrSlot &ts = wslots[tgt];
                    h_a16 = ts.a16, h_b8 = ts.b8, h_amis = ts.a_mis, h_bmis = ts.b_mis, h_last8 = ts.last8;
                    h_c = ts.sc, h_lim = ts.slim;
                    h_row = ts.srow + 128 * boff;
                    h_step = 128 * power;
                    g_boff = (uint32_t)boff;
                    serving = true;
                    __syncwarp();
                }
                const int row = h_row + 16 * (int)l;  // rows of `a` are 16-byte aligned here
                int r_l = 0x7FFFFFFF;                 // "this lane does not end the slide"
                if (serving) {
                    if (row < h_lim) {
                        const uint32_t ia = ((uint32_t)h_amis + (uint32_t)row) >> 4;
                        const int pbs = h_bmis + row + h_c;  // >= 0: rows of a slide lie past the diagonal's first row
                        const int j = pbs >> 3;
                        const uint4 A = __ldg(h_a16 + ia);
                        uint2 v0 = make_uint2(0u, 0u), v1 = v0, v2 = v0;
                        if (j <= h_last8) v0 = __ldg(h_b8 + j);
                        if (j + 1 <= h_last8) v1 = __ldg(h_b8 + j + 1);
                        if (j + 2 <= h_last8) v2 = __ldg(h_b8 + j + 2);
                        uint32_t x0, x1, x2, x3;
                        xor16(A, v0, v1, v2, (uint32_t)pbs & 7u, x0, x1, x2, x3);
                        if ((x0 | x1 | x2 | x3) != 0u)
                            r_l = min(row + (int)first_nonzero_byte(x0, x1, x2, x3), h_lim);
                        else if (row + 16 >= h_lim)
                            r_l = h_lim;
                    } else {
                        r_l = h_lim;  // the slide ended before this lane's chunk
                    }
                }
                // rows grow with (block, lane), so the lane that ends the slide is the one with the smallest result
                const int rr = __reduce_min_sync(gmask, r_l);
                const bool ended = rr != 0x7FFFFFFF;
                if (ended) {
                    // The next level resumes on the neighbouring diagonals just past row rr: its probes and the first
                    // step of its slide read the lines that follow.  Ask L2 for them now (one 128-byte line of each
                    // string from two lanes of every serving octet; all eight lanes fetched 1.8x the algorithmic bytes), so
                    // that they arrive while the round in between is computed instead of costing that step a DRAM round
                    // trip.
                    if (serving && l < 2u) {  // two lines per serving octet: 256 .. 1024 bytes ahead of rr
                        const int prow = rr + 128 * (int)(2u * g_boff + l);
                        if (prow < h_lim + 128) {
                            prefetch_l2_line((const uint8_t *)h_a16 + h_amis + min(prow, h_lim - 1));
                            prefetch_l2_line((const uint8_t *)h_b8 + h_bmis + min(prow, h_lim - 1) + h_c);
                        }
                    }
                    if (in_slide) {
                        if (l == q) v = rr;
                        P &= P - 1u;
                        in_slide = false;
                    }
                    serving = false;  // until the next assignment
                } else {
                    h_row += h_step;
                    if (in_slide) own_row = h_row;
                }
                // Whenever a slide ends its octets are re-assigned at once: they either start their next pending
                // diagonal or join the slides that go on.  (Re-assigning only when a new slide starts left the helpers
                // of finished slides idle: 1.62 ms instead of 1.11 ms on 256 Ki pairs of 4096 bytes.)
                assign = __any_sync(FULL, ended);
            }
        }

        // ---- D: publish the round, test for the end of the pair, advance the level ----------------------------------
        if (valid) cur[s] = v;
        const bool hit = valid && c == diff && v >= m;
        const uint32_t hm = (__ballot_sync(FULL, hit) >> osh) & 0xFFu;
        if (have) {
            if (hm) {
                if (l == 0) out[pair] = (uint32_t)e;
                have = false;
            } else {
                cbase += 8;
                if (cbase > hi) {
                    e++;
                    if (e > max_k) {
                        if (l == 0) out[pair] = 0xFFFFFFFFu;
                        have = false;
                    } else {
                        const int lo = max(max(-e, -m), diff - (max_k - e));
                        hi = min(e, diff + (max_k - e));
                        cbase = lo;
                    }
                }
            }
        }
        __syncwarp();
    }
}


}  // namespace

// Largest min(k, max_len) this kernel takes (the level arrays live in shared memory: 2 (2 k + 3) words per octet)
#define TA_FR_MAX_K 64u

bool ta_fr_can_handle(uint32_t k, ta_costs c, uint32_t max_len) {
    if (!(c.mismatch == 1 && c.gap == 1 && c.start_gap == 0 && c.transpose <= 1)) return false;
    const uint32_t kk = k < max_len ? k : max_len;
    return kk <= TA_FR_MAX_K;
}

// Is diagonal extension expected to beat the bit-parallel band kernels?  Filling the band costs ~30 instructions per
// column per pair whatever the data; this kernel costs ~20 warp instructions per round (about (k + 1)^2 / 10 rounds
// for a pair at distance k) plus ~10 per 128 bytes.  Measured cross-over: profiles/README.md (round 2).
bool ta_fr_preferred(uint32_t k, ta_costs c, uint32_t max_len) {
    static const int force = getenv("TA_FR") ? atoi(getenv("TA_FR")) : -1;  // 0 = never, 1 = whenever possible
    if (!ta_fr_can_handle(k, c, max_len)) return false;
    if (force == 0) return false;
    if (force == 1) return true;
    const uint32_t kk = k < max_len ? k : max_len;
    return (uint64_t)max_len >= 1024u && (uint64_t)max_len >= 4ull * kk * kk;
}

int ta_launch_lev_fr(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                     size_t n, const uint32_t *idx, uint32_t k, ta_costs costs, uint32_t max_len, uint32_t *out,
                     cudaStream_t st) {
    if (n == 0) return TA_OK;
    const uint32_t kk = k < max_len ? k : max_len;
    static const int env_per_sm = getenv("TA_FR_CTAS") ? atoi(getenv("TA_FR_CTAS")) : 0;
    const int nt = 128;
    // 2 kk + 3 slots, rounded so that an octet's two arrays are 16-byte aligned and successive octets start 8 banks apart
    uint32_t slots_pad = (2u * kk + 3u + 3u) & ~3u;
    while ((2u * slots_pad) % 32u != 8u) slots_pad += 4u;
    const size_t smem = (size_t)(nt / 8) * (2u * slots_pad * sizeof(int) + sizeof(FrSlot));
    auto kern = costs.transpose ? lev_fr_kernel<true> : lev_fr_kernel<false>;
    // persistent octets: exactly as many CTAs as are resident at once (a second, partial wave would leave SMs idle)
    int per_sm = 0;
    TA_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nt, smem));
    if (per_sm < 1) per_sm = 1;
    if (env_per_sm > 0) per_sm = env_per_sm;
    const size_t octets = (size_t)nt / 8;
    const unsigned blocks = (unsigned)std::min<size_t>((n + octets - 1) / octets, (size_t)ctx->sm_count * per_sm);
    kern<<<blocks, nt, smem, st>>>(a, a_off, b, b_off, idx, n, k, slots_pad, out);
    ctx->launches++;
    TA_CUDA(ctx, cudaGetLastError());
    return TA_OK;
}
