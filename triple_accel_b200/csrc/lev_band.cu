// lev_band.cu -- general k-banded anti-diagonal edit-distance DP for sm_100a (all EditCosts, any k).
//
// Contract (reference src/levenshtein.rs:376-545 scalar, 714-827 SIMD entry): for each pair, the weighted edit
// distance d under EditCosts {mismatch, gap, start_gap (affine), transpose (restricted Damerau)} if d <= k,
// else TA_NONE.  `max_k`/`unit_k` clamps and the early None follow src/levenshtein.rs:400-430.
//
// Design (not a port of the reference's AVX2 lanes): a group of G lanes (G = 2..32, several pairs per warp for
// narrow bands) owns one pair.  Lane t holds C cells of the current anti-diagonal s = i + j in registers; cell
// ci = t*C + c sits on diagonal d = dlo + 2*ci + p with p = (s - dlo) & 1, so consecutive anti-diagonals
// alternate between "even" and "odd" diagonals and only ceil(W/2) cells are live per step (W = band width).
// The band is Ukkonen's: only diagonals a path of cost <= max_k can touch,
//     [-e, diff + e],  e = (max_k - 2*start_gap - diff*gap) / (2*gap),
// which is about half the reference's [-unit_k, unit_k] band and gives identical answers for d <= max_k.
// Per step each lane needs one neighbour value: on p == 0 steps the "left" cell (diagonal d-1) comes from lane
// t-1 (__shfl_up_sync), on p == 1 steps the "up" cell (d+1) from lane t+1 (__shfl_down_sync).  Affine gaps are
// carried as the sender-side pre-minimised values outH = min(D + open, H + gap), outV likewise, so a step costs
// one shuffle whatever the cost model; the transposition test needs the neighbours' match flags, which ride in
// bit 0 of the shuffled word.  Strings are staged into shared memory with 16-byte cp.async (LDGSTS) vectors.
#include <stdlib.h>

#include <algorithm>

#include "lev_band_common.cuh"
#include "ta_common.cuh"

namespace {

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ uint32_t umin3(uint32_t x, uint32_t y, uint32_t z) { return min(min(x, y), z); }

template <int G, int C, bool AFFINE, bool TRANS, bool SMEM, bool TRACE>
__global__ void __launch_bounds__(128) lev_band_kernel(const BandArgs args) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int GROUPS = 128 / G;
    const int grp = threadIdx.x / G;
    const int t = threadIdx.x % G;  // lane within the group
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));

    const size_t w = (size_t)blockIdx.x * GROUPS + grp;
    if (w >= args.n) return;  // whole group leaves together (shuffles below use the group mask)
    const size_t pair = args.idx ? (size_t)args.idx[w] : args.pair_base + w;

    // ---- per-pair setup (every lane computes the same scalars) --------------------------------------------------
    const uint64_t a0 = args.a_off[pair], a1 = args.a_off[pair + 1];
    const uint64_t b0 = args.b_off[pair], b1 = args.b_off[pair + 1];
    const bool swap = (a1 - a0) > (b1 - b0);  // src/levenshtein.rs:386 -- "a" is the shorter string
    const uint8_t *ga = swap ? args.b + b0 : args.a + a0;
    const uint8_t *gb = swap ? args.a + a0 : args.b + b0;
    const int m = (int)(swap ? (b1 - b0) : (a1 - a0));
    const int n = (int)(swap ? (a1 - a0) : (b1 - b0));
    const uint32_t mism = args.mism, gap = args.gap, sgap = args.sgap, tcost = args.tcost;
    const uint32_t diff = (uint32_t)(n - m);

    const BandInfo bi = band_info(m, n, args.k, mism, gap, sgap, TRANS);
    const uint32_t max_k = bi.max_k;
    if (bi.none) {  // src/levenshtein.rs:428-430
        if (t == 0) args.out[pair] = TA_NONE;
        return;
    }
    if (m == 0) {  // D(0, n) = n*gap + start_gap
        if (t == 0) {
            const uint32_t d = (uint32_t)n * gap + (n ? sgap : 0u);
            args.out[pair] = d <= max_k ? d : TA_NONE;
        }
        return;
    }
    const int dlo = bi.dlo;  // host guarantees bi.W <= 2*G*C

    // ---- stage both strings into shared memory (16-byte cp.async vectors on the aligned-down addresses) --------
    const uint8_t *sa, *sb;
    if (SMEM) {
        uint8_t *slot_a = smem + (size_t)grp * 2 * args.slot;
        uint8_t *slot_b = slot_a + args.slot;
        const uintptr_t ua = (uintptr_t)ga, ub = (uintptr_t)gb;
        const uint8_t *ga16 = (const uint8_t *)(ua & ~(uintptr_t)15), *gb16 = (const uint8_t *)(ub & ~(uintptr_t)15);
        const int sha = (int)(ua & 15), shb = (int)(ub & 15);
        const int na = (sha + m + 15) >> 4, nb = (shb + n + 15) >> 4;
        for (int q = t; q < na; q += G) cp_async16(slot_a + 16 * q, ga16 + 16 * q);
        for (int q = t; q < nb; q += G) cp_async16(slot_b + 16 * q, gb16 + 16 * q);
        cp_async_wait_all();
        __syncwarp(gmask);
        sa = slot_a + sha;
        sb = slot_b + shb;
    } else {
        sa = ga;
        sb = gb;
    }
    auto A = [&](int i) -> uint32_t { return sa[__vimin_s32_relu(i, m - 1)]; };  // clamped: out-of-matrix cells
    auto B = [&](int j) -> uint32_t { return sb[__vimin_s32_relu(j, n - 1)]; };  // never reach the result

    const uint32_t open = sgap + gap;

    // ---- DP state --------------------------------------------------------------------------------------------
    uint32_t D1[C], D2[C], D3[C], D4[C];  // this lane's cells on anti-diagonals s-1 .. s-4
    uint32_t oH[C], oV[C];                // pre-minimised gap candidates offered to the right / lower neighbour
    uint32_t mt[C];                       // match flag of the s-1 cells (TRANS)
    uint32_t ca[C], cb[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        D1[c] = D2[c] = D3[c] = D4[c] = TA_INF;
        oH[c] = oV[c] = TA_INF;
        mt[c] = 0;
    }

    // first anti-diagonal processed has (s - dlo) even
    int s = -((-dlo) & 1);
    // coordinates of cell c == 0 on the upcoming p == 0 step: i = (s - d)/2, j = (s + d)/2, d = dlo + 2*t*C
    int i0 = (s - dlo) / 2 - t * C;
    int j0 = (s + dlo) / 2 + t * C;  // (s + dlo) is even, may be negative: exact division
#pragma unroll
    for (int c = 0; c < C; c++) cb[c] = B(j0 + c - 1);
#pragma unroll
    for (int c = 0; c < C; c++) ca[c] = A(i0 - c - 1 - 1);  // the p == 0 step below shifts these by one first

    const int s_end = m + n;
    // steps with s <= s_bnd may touch the first row / column and need the boundary override
    const int s_bnd = max(-dlo, dlo + 2 * G * C);

    uint8_t *trow = TRACE ? args.trace + w * args.trace_stride + (size_t)t * C : nullptr;  // advances G*C per step
    auto step = [&](const int p, const bool boundary) {
        // p == 0: i advanced by one since the previous step; p == 1: j advanced by one
        uint32_t rH, rV;
        if (p == 0) {
#pragma unroll
            for (int c = C - 1; c > 0; c--) ca[c] = ca[c - 1];
            ca[0] = A(i0 - 1);
            uint32_t send = AFFINE ? oH[C - 1] : oV[C - 1];
            if (TRANS) send = (min(send, TA_INF) << 1) | mt[C - 1];
            rH = __shfl_up_sync(gmask, send, 1, G);
            if (t == 0) rH = TRANS ? (TA_INF << 1) : TA_INF;
            rV = 0;
        } else {
#pragma unroll
            for (int c = 0; c < C - 1; c++) cb[c] = cb[c + 1];
            cb[C - 1] = B(j0 + C - 1 - 1);
            uint32_t send = oV[0];
            if (TRANS) send = (min(send, TA_INF) << 1) | mt[0];
            rV = __shfl_down_sync(gmask, send, 1, G);
            if (t == G - 1) rV = TRANS ? (TA_INF << 1) : TA_INF;
            rH = 0;
        }
        uint32_t rflag = 0;
        if (TRANS) {
            if (p == 0) {
                rflag = rH & 1;
                rH >>= 1;
            } else {
                rflag = rV & 1;
                rV >>= 1;
            }
        }
        uint32_t nD[C], nH[C], nV[C], nM[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            uint32_t h, v, mL = 0, mU = 0;
            if (p == 0) {
                h = (c == 0) ? rH : (AFFINE ? oH[c - 1] : oV[c - 1]);
                v = oV[c];
                if (TRANS) {
                    mL = (c == 0) ? rflag : mt[c - 1];
                    mU = mt[c];
                }
            } else {
                h = AFFINE ? oH[c] : oV[c];
                v = (c == C - 1) ? rV : oV[c + 1];
                if (TRANS) {
                    mL = mt[c];
                    mU = (c == C - 1) ? rflag : mt[c + 1];
                }
            }
            const uint32_t eq = ca[c] == cb[c];
            uint32_t d, arg = 0;
            if (TRACE) {
                // the scalar routine's decision order (src/levenshtein.rs:493-532): substitution, a-gap if <,
                // b-gap if <, transposition if <=
                d = D2[c] + (eq ? 0u : mism);
                if (h < d) {
                    d = h;
                    arg = 1;
                }
                if (v < d) {
                    d = v;
                    arg = 2;
                }
                if (TRANS && (mL & mU)) {
                    const uint32_t tr = D4[c] + tcost;
                    if (tr <= d) {
                        d = tr;
                        arg = 3;
                    }
                }
            } else {
                d = umin3(D2[c] + (eq ? 0u : mism), h, v);
                if (TRANS) {
                    const uint32_t tr = D4[c] + tcost;
                    if (mL & mU) d = min(d, tr);
                }
            }
            uint32_t hh = h, vv = v;
            if (boundary) {
                const int i = i0 - c, j = j0 + c;
                const bool bi = (i == 0) & (j >= 0), bj = (j == 0) & (i >= 0);
                if (bi | bj) {
                    const int q = bi ? j : i;
                    d = (uint32_t)q * gap + (q > 0 ? sgap : 0u);
                    hh = vv = TA_INF;
                    arg = bi ? 1u : 2u;  // first row: a-gaps (:450-456); first column: b-gaps
                }
            }
            if (TRACE) trow[c] = (uint8_t)arg;
            nD[c] = d;
            if (AFFINE) {
                nH[c] = min(d + open, hh + gap);
                nV[c] = min(d + open, vv + gap);
            } else {
                nV[c] = d + gap;
            }
            nM[c] = eq;
        }
#pragma unroll
        for (int c = 0; c < C; c++) {
            if (TRANS) {
                D4[c] = D3[c];
                D3[c] = D2[c];
                mt[c] = nM[c];
            }
            D2[c] = D1[c];
            D1[c] = nD[c];
            if (AFFINE) oH[c] = nH[c];
            oV[c] = nV[c];
        }
        if (TRACE) trow += G * C;
    };

    // phase 1: anti-diagonals that can contain first-row / first-column cells
    for (; s <= s_end && s <= s_bnd; s += 2) {
        step(0, true);
        j0 += 1;
        step(1, true);
        i0 += 1;
    }
    // phase 2: interior
    for (; s <= s_end; s += 2) {
        step(0, false);
        j0 += 1;
        step(1, false);
        i0 += 1;
    }

    // the loop ran pairs (p = 0 at s, p = 1 at s + 1); the cell (m, n) was produced on anti-diagonal m + n
    const int pf = (s_end - dlo) & 1;
    const int cif = ((int)diff - dlo - pf) >> 1;
    const int tf = cif / C, cf = cif % C;
    uint32_t val = 0;
#pragma unroll
    for (int c = 0; c < C; c++) {
        // after the final pair of steps D1 = p==1 output (anti-diagonal s-1), D2 = p==0 output (s-2)
        const uint32_t x = pf ? D1[c] : D2[c];
        if (c == cf) val = x;
    }
    val = __shfl_sync(gmask, val, tf, G);
    if (t == 0) args.out[pair] = val <= max_k ? val : TA_NONE;
}

// ---------------------------------------------------------------------------------------------------------------
// lev_wide_kernel: one BLOCK per pair for bands wider than 1024 diagonals (levenshtein()/rdamerau() on long
// strings, late rounds of the exponential search).  Same anti-diagonal recurrence and band as above, but the
// anti-diagonals live in shared memory (or, for very wide bands, in a per-block global workspace) and the block
// advances one anti-diagonal per __syncthreads().
struct WideArgs {
    BandArgs b;
    uint32_t wc;         // cells per anti-diagonal the buffers are sized for
    uint32_t *workspace;  // null: use dynamic shared memory; else gridDim.x * 10 * wc words
};

__global__ void __launch_bounds__(256) lev_wide_kernel(const WideArgs wa) {
    extern __shared__ __align__(16) uint8_t smem[];
    const BandArgs &args = wa.b;
    uint32_t *base = wa.workspace ? wa.workspace + (size_t)blockIdx.x * 10 * wa.wc : (uint32_t *)smem;
    const uint32_t wc = wa.wc;
    for (size_t w = blockIdx.x; w < args.n; w += gridDim.x) {
        const size_t pair = args.idx ? (size_t)args.idx[w] : args.pair_base + w;
        const uint64_t a0 = args.a_off[pair], a1 = args.a_off[pair + 1];
        const uint64_t b0 = args.b_off[pair], b1 = args.b_off[pair + 1];
        const bool swap = (a1 - a0) > (b1 - b0);
        const uint8_t *sa = swap ? args.b + b0 : args.a + a0;
        const uint8_t *sb = swap ? args.a + a0 : args.b + b0;
        const int m = (int)(swap ? (b1 - b0) : (a1 - a0));
        const int n = (int)(swap ? (a1 - a0) : (b1 - b0));
        const uint32_t mism = args.mism, gap = args.gap, sgap = args.sgap, tcost = args.tcost;
        const uint32_t diff = (uint32_t)(n - m);
        const bool trans = tcost != 0;
        const BandInfo bi = band_info(m, n, args.k, mism, gap, sgap, trans);
        const uint32_t max_k = bi.max_k;
        __syncthreads();  // previous pair's buffers are no longer read
        if (bi.none) {
            if (threadIdx.x == 0) args.out[pair] = TA_NONE;
            continue;
        }
        if (m == 0) {
            if (threadIdx.x == 0) {
                const uint32_t d = (uint32_t)n * gap + (n ? sgap : 0u);
                args.out[pair] = d <= max_k ? d : TA_NONE;
            }
            continue;
        }
        const int dlo = bi.dlo;
        const int W = bi.W;
        const int cells = (W + 1) / 2 + 1;  // <= wc (host)
        uint32_t *D[4] = {base, base + wc, base + 2 * wc, base + 3 * wc};  // D[0] = s-1, D[1] = s-2, ...
        uint32_t *oH[2] = {base + 4 * wc, base + 5 * wc}, *oV[2] = {base + 6 * wc, base + 7 * wc};
        uint32_t *mt[2] = {base + 8 * wc, base + 9 * wc};
        for (int ci = threadIdx.x; ci < cells; ci += blockDim.x) {
            D[0][ci] = D[1][ci] = D[2][ci] = D[3][ci] = TA_INF;
            oH[0][ci] = oV[0][ci] = TA_INF;
            mt[0][ci] = 0;
        }
        __syncthreads();
        const uint32_t open = sgap + gap;
        const int s_end = m + n;
        const int s0 = -((-dlo) & 1);
        int cur = 0;  // oH/oV/mt[cur] hold step s-1
        for (int s = s0; s <= s_end + 1; s++) {
            const int p = (s - dlo) & 1;
            uint32_t *Dn = trans ? D[3] : D[1];  // overwritten in place (each cell only reads its own old slot)
            for (int ci = threadIdx.x; ci < cells; ci += blockDim.x) {
                const int d = dlo + 2 * ci + p;
                const int i = (s - d) >> 1, j = (s + d) >> 1;  // (s - d) is even
                const int lh = p ? ci : ci - 1, lv = p ? ci + 1 : ci;
                uint32_t h = (lh >= 0 && lh < cells) ? oH[cur][lh] : TA_INF;
                uint32_t v = (lv >= 0 && lv < cells) ? oV[cur][lv] : TA_INF;
                const uint32_t ca = sa[__vimin_s32_relu(i - 1, m - 1)], cb = sb[__vimin_s32_relu(j - 1, n - 1)];
                const uint32_t eq = ca == cb;
                uint32_t val = umin3(D[1][ci] + (eq ? 0u : mism), h, v);
                if (trans) {
                    const uint32_t mL = (lh >= 0 && lh < cells) ? mt[cur][lh] : 0u;
                    const uint32_t mU = (lv >= 0 && lv < cells) ? mt[cur][lv] : 0u;
                    if (mL & mU) val = min(val, D[3][ci] + tcost);
                    mt[cur ^ 1][ci] = eq;
                }
                const bool bi = (i == 0) & (j >= 0), bj = (j == 0) & (i >= 0);
                if (bi | bj) {
                    const int q = bi ? j : i;
                    val = (uint32_t)q * gap + (q > 0 ? sgap : 0u);
                    h = v = TA_INF;
                }
                val = min(val, TA_INF);
                Dn[ci] = val;
                oH[cur ^ 1][ci] = min(min(val + open, h + gap), TA_INF);
                oV[cur ^ 1][ci] = min(min(val + open, v + gap), TA_INF);
            }
            // rotate: new -> s-1
            if (trans) {
                uint32_t *t3 = D[3];
                D[3] = D[2];
                D[2] = D[1];
                D[1] = D[0];
                D[0] = t3;
            } else {
                uint32_t *t1 = D[1];
                D[1] = D[0];
                D[0] = t1;
            }
            cur ^= 1;
            __syncthreads();
            if (s == s_end) {
                const int cif = ((int)diff - dlo - p) >> 1;
                if (threadIdx.x == 0) {
                    const uint32_t val = D[0][cif];
                    args.out[pair] = val <= max_k ? val : TA_NONE;
                }
                break;
            }
        }
    }
}

int launch_wide(ta_ctx *ctx, const BandArgs &args, uint32_t W, cudaStream_t st) {
    WideArgs wa;
    wa.b = args;
    wa.wc = (W + 1) / 2 + 2;
    const size_t bytes = (size_t)10 * wa.wc * sizeof(uint32_t);
    unsigned blocks = (unsigned)std::min<size_t>(args.n, (size_t)ctx->sm_count * 2);
    if (bytes <= (size_t)ctx->smem_optin - 1024) {
        wa.workspace = nullptr;
        TA_CUDA(ctx, cudaFuncSetAttribute(lev_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        lev_wide_kernel<<<blocks, 256, bytes, st>>>(wa);
    } else {
        int rc = ta_dev_reserve(ctx, ctx->d_work[3], bytes * blocks);
        if (rc != TA_OK) return rc;
        wa.workspace = (uint32_t *)ctx->d_work[3].p;
        lev_wide_kernel<<<blocks, 256, 0, st>>>(wa);
    }
    ctx->launches++;
    TA_CUDA(ctx, cudaGetLastError());
    return TA_OK;
}

template <int G, int C, bool AFFINE, bool TRANS, bool TRACE = false>
int launch_gc(ta_ctx *ctx, const BandArgs &args0, uint32_t max_len, cudaStream_t st) {
    BandArgs args = args0;
    constexpr int GROUPS = 128 / G;
    const size_t blocks = (args.n + GROUPS - 1) / GROUPS;
    const uint32_t slot = ((max_len + 15u) & ~15u) + 32u;  // + alignment slack of the aligned-down copy
    const size_t smem = (size_t)GROUPS * 2 * slot;
    if (smem <= 96 * 1024) {
        args.slot = slot;
        auto kern = lev_band_kernel<G, C, AFFINE, TRANS, true, TRACE>;
        if (smem > 48 * 1024)
            TA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, 128, smem, st>>>(args);
    } else {
        args.slot = 0;
        lev_band_kernel<G, C, AFFINE, TRANS, false, TRACE><<<(unsigned)blocks, 128, 0, st>>>(args);
    }
    ctx->launches++;
    TA_CUDA(ctx, cudaGetLastError());
    return TA_OK;
}

template <bool AFFINE, bool TRANS>
int launch_w(ta_ctx *ctx, const BandArgs &args, uint32_t W, uint32_t max_len, cudaStream_t st) {
    // four cells per lane wherever the band allows it: the shuffle, the character load and the loop bookkeeping are
    // shared by the lane's cells (measured on cfg 2 with TA_FORCE_BAND=1: 1.08 ms at one cell per lane, 0.82 ms at
    // two, 0.70 ms at four; eight cells on one or two lanes lose to shared-memory bank conflicts)
    if (W <= 16) return launch_gc<2, 4, AFFINE, TRANS>(ctx, args, max_len, st);
    if (W <= 32) return launch_gc<4, 4, AFFINE, TRANS>(ctx, args, max_len, st);
    if (W <= 64) return launch_gc<8, 4, AFFINE, TRANS>(ctx, args, max_len, st);
    if (W <= 128) return launch_gc<16, 4, AFFINE, TRANS>(ctx, args, max_len, st);
    if (W <= 256) return launch_gc<32, 4, AFFINE, TRANS>(ctx, args, max_len, st);
    if (W <= 512) return launch_gc<32, 8, AFFINE, TRANS>(ctx, args, max_len, st);
    if (W <= 1024) return launch_gc<32, 16, AFFINE, TRANS>(ctx, args, max_len, st);
    return launch_wide(ctx, args, W, st);
}

// ---------------------------------------------------------------------------------------------------------------
// Traceback walk (reference src/levenshtein.rs:547-606): one thread per pair follows the per-cell decisions stored
// by the TRACE variant of lev_band_kernel from (m, n) back to (0, 0) and run-length encodes the edits.  Pass 1
// (edits == nullptr) counts the runs, pass 2 writes them in forward order at edit_off[w].
struct WalkArgs {
    BandArgs b;
    uint32_t wc;               // cells per anti-diagonal in the trace (G * C of the kernel that wrote it)
    uint32_t *counts;          // [n] runs per work item
    const uint64_t *edit_off;  // [n] first output slot per work item (pass 2)
    ta_edit *edits;            // pass 2 output, or nullptr for pass 1
};

__global__ void __launch_bounds__(128) trace_walk_kernel(const WalkArgs wa) {
    const BandArgs &args = wa.b;
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= args.n) return;
    const size_t pair = args.idx ? (size_t)args.idx[w] : args.pair_base + w;
    const uint64_t a0 = args.a_off[pair], a1 = args.a_off[pair + 1];
    const uint64_t b0 = args.b_off[pair], b1 = args.b_off[pair + 1];
    const bool swap = (a1 - a0) > (b1 - b0);
    const uint8_t *sa = swap ? args.b + b0 : args.a + a0;
    const uint8_t *sb = swap ? args.a + a0 : args.b + b0;
    const int m = (int)(swap ? (b1 - b0) : (a1 - a0));
    const int n = (int)(swap ? (a1 - a0) : (b1 - b0));
    const bool writing = wa.edits != nullptr;
    uint32_t runs = 0;
    if (args.out[pair] != TA_NONE && (m > 0 || n > 0)) {
        const uint32_t total = writing ? wa.counts[w] : 0u;
        ta_edit *dst = writing ? wa.edits + wa.edit_off[w] : nullptr;
        uint32_t cur = 0xffffffffu, cnt = 0;
        auto push = [&](uint32_t e) {  // src/levenshtein.rs:598-602 (reversed order: the walk goes backwards)
            if (e == cur) {
                cnt++;
                return;
            }
            if (cnt) {
                if (writing) dst[total - 1 - runs] = ta_edit{cur, cnt};
                runs++;
            }
            cur = e;
            cnt = 1;
        };
        const uint32_t e_agap = swap ? 3u : 2u, e_bgap = swap ? 2u : 3u;  // AGap = 2, BGap = 3 (src/lib.rs:147-154)
        if (m == 0) {
            for (int j = n; j > 0; j--) push(e_agap);  // first row: every step is an a-gap (:450-456, 574-581)
        } else {
            const BandInfo bi = band_info(m, n, args.k, args.mism, args.gap, args.sgap, args.tcost != 0);
            const int dlo = bi.dlo, s0 = -((-dlo) & 1);
            const uint8_t *tr = args.trace + w * args.trace_stride;
            int i = m, j = n;
            while (i > 0 || j > 0) {
                const int s = i + j, d = j - i;
                const int p = (s - dlo) & 1;
                const int ci = (d - dlo - p) >> 1;
                const uint32_t arg = tr[(size_t)(s - s0) * wa.wc + ci];
                if (arg == 0) {
                    i--;
                    j--;
                    push(sa[i] == sb[j] ? 0u : 1u);  // Match / Mismatch
                } else if (arg == 1) {
                    j--;
                    push(e_agap);
                } else if (arg == 2) {
                    i--;
                    push(e_bgap);
                } else {
                    i -= 2;
                    j -= 2;
                    push(4u);  // Transpose
                }
            }
        }
        if (cnt) {
            if (writing) dst[total - 1 - runs] = ta_edit{cur, cnt};
            runs++;
        }
    }
    if (!writing) wa.counts[w] = runs;
}

template <bool TRANS>
int launch_trace_w(ta_ctx *ctx, const BandArgs &args, uint32_t W, uint32_t max_len, cudaStream_t st, uint32_t *wc) {
    // affine formulas are valid for start_gap == 0 too: the TRACE variants are only instantiated with AFFINE = true
    if (W <= 16) return *wc = 8, launch_gc<2, 4, true, TRANS, true>(ctx, args, max_len, st);
    if (W <= 32) return *wc = 16, launch_gc<4, 4, true, TRANS, true>(ctx, args, max_len, st);
    if (W <= 64) return *wc = 32, launch_gc<8, 4, true, TRANS, true>(ctx, args, max_len, st);
    if (W <= 128) return *wc = 64, launch_gc<16, 4, true, TRANS, true>(ctx, args, max_len, st);
    if (W <= 256) return *wc = 128, launch_gc<32, 4, true, TRANS, true>(ctx, args, max_len, st);
    if (W <= 512) return *wc = 256, launch_gc<32, 8, true, TRANS, true>(ctx, args, max_len, st);
    if (W <= 1024) return *wc = 512, launch_gc<32, 16, true, TRANS, true>(ctx, args, max_len, st);
    return TA_ERR_TOO_LARGE;
}

}  // namespace

uint32_t ta_band_width_bound(uint32_t k, ta_costs c, uint32_t max_len) {
    // max_k <= min(k, m*mismatch + diff*gap + start_gap) <= min(k, max_len * max(mismatch, gap) + start_gap);
    // W = diff + 2e + 1 <= (max_k - 2*start_gap)/gap + 1  (+2 with transpositions)
    const uint64_t ub = (uint64_t)max_len * (c.mismatch > c.gap ? c.mismatch : c.gap) + c.start_gap;
    const uint64_t kk = k < ub ? k : ub;
    const uint64_t unit = (kk > c.start_gap ? kk - c.start_gap : 0) / c.gap;
    uint64_t W = unit + 1 + (c.transpose ? 2 : 0);
    const uint64_t full = 2ull * max_len + 1 + (c.transpose ? 2 : 0);  // never wider than the whole matrix
    if (W > full) W = full;
    return (uint32_t)(W > 0xFFFFFFFFull ? 0xFFFFFFFFull : W);
}

int ta_launch_lev_band(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                       size_t n, const uint32_t *idx, uint32_t k, ta_costs costs, uint32_t max_len, uint32_t *out,
                       cudaStream_t st) {
    if (n == 0) return TA_OK;
    BandArgs args;
    args.a = a, args.a_off = a_off, args.b = b, args.b_off = b_off, args.idx = idx, args.n = n, args.k = k;
    args.mism = costs.mismatch, args.gap = costs.gap, args.sgap = costs.start_gap, args.tcost = costs.transpose;
    args.slot = 0, args.out = out, args.trace = nullptr, args.trace_stride = 0, args.pair_base = 0;
    const uint32_t W = ta_band_width_bound(k, costs, max_len);
    // narrow bands, large batches: one thread per pair with packed u16 cells (lev_diag16.cu)
    if (ta_diag16_can_handle(n, k, costs, max_len, W)) return ta_launch_lev_diag16(ctx, args, costs, W, st);
    const bool affine = costs.start_gap != 0, trans = costs.transpose != 0;
    if (affine && trans) return launch_w<true, true>(ctx, args, W, max_len, st);
    if (affine) return launch_w<true, false>(ctx, args, W, max_len, st);
    if (trans) return launch_w<false, true>(ctx, args, W, max_len, st);
    return launch_w<false, false>(ctx, args, W, max_len, st);
}

// cells per anti-diagonal (G * C) the trace kernel will use for band width W; 0 if the band is too wide
uint32_t ta_trace_cells(uint32_t W) {
    if (W <= 16) return 8;
    if (W <= 32) return 16;
    if (W <= 64) return 32;
    if (W <= 128) return 64;
    if (W <= 256) return 128;
    if (W <= 512) return 256;
    if (W <= 1024) return 512;
    return 0;
}

// Distances + per-cell decisions for work items [0, n) (pairs pair_base + w, or idx[w]); trace needs
// n * trace_stride bytes with trace_stride >= (2 * max_len + 3) * ta_trace_cells(W).
int ta_launch_lev_band_trace(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                             const uint64_t *b_off, size_t n, const uint32_t *idx, size_t pair_base, uint32_t k,
                             ta_costs costs, uint32_t max_len, uint32_t *out, uint8_t *trace, size_t trace_stride,
                             cudaStream_t st) {
    if (n == 0) return TA_OK;
    BandArgs args;
    args.a = a, args.a_off = a_off, args.b = b, args.b_off = b_off, args.idx = idx, args.pair_base = pair_base;
    args.n = n, args.k = k;
    args.mism = costs.mismatch, args.gap = costs.gap, args.sgap = costs.start_gap, args.tcost = costs.transpose;
    args.slot = 0, args.out = out, args.trace = trace, args.trace_stride = trace_stride;
    const uint32_t W = ta_band_width_bound(k, costs, max_len);
    uint32_t wc = 0;
    return costs.transpose ? launch_trace_w<true>(ctx, args, W, max_len, st, &wc)
                           : launch_trace_w<false>(ctx, args, W, max_len, st, &wc);
}

// pass 1 (edits == nullptr): counts[w] = number of runs; pass 2: writes them at edits[edit_off[w] ..)
int ta_launch_trace_walk(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                         size_t n, const uint32_t *idx, size_t pair_base, uint32_t k, ta_costs costs, uint32_t max_len,
                         const uint32_t *out, const uint8_t *trace, size_t trace_stride, uint32_t *counts,
                         const uint64_t *edit_off, ta_edit *edits, cudaStream_t st) {
    if (n == 0) return TA_OK;
    WalkArgs wa;
    wa.b.a = a, wa.b.a_off = a_off, wa.b.b = b, wa.b.b_off = b_off, wa.b.idx = idx, wa.b.pair_base = pair_base;
    wa.b.n = n, wa.b.k = k;
    wa.b.mism = costs.mismatch, wa.b.gap = costs.gap, wa.b.sgap = costs.start_gap, wa.b.tcost = costs.transpose;
    wa.b.slot = 0, wa.b.out = const_cast<uint32_t *>(out), wa.b.trace = const_cast<uint8_t *>(trace);
    wa.b.trace_stride = trace_stride;
    wa.wc = ta_trace_cells(ta_band_width_bound(k, costs, max_len));
    wa.counts = counts, wa.edit_off = edit_off, wa.edits = edits;
    trace_walk_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(wa);
    ctx->launches++;
    TA_CUDA(ctx, cudaGetLastError());
    return TA_OK;
}
