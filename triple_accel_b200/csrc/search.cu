// search.cu -- needle-in-haystack Levenshtein search for sm_100a.
//
// Contract: levenshtein_search_simd_with_opts (reference src/levenshtein.rs:1911-2155) bit-exact with the scalar
// levenshtein_search_naive_with_opts (src/levenshtein.rs:1589-1838): every end position whose best semi-global
// alignment of the whole needle costs <= k is reported as Match{start = end - length, end, k = cost}, where
// `length` follows the reference's tie-breaking rules cell by cell (src/levenshtein.rs:1726-1779, including the
// comparison against length2[j-1] at :1756).  SearchType::Best is the running-minimum filter of :1792-1796
// followed by the overlap / minimum post-pass of :1812-1835; it is applied to the (sparse) hit list.
//
// Kernels:
//  * search_exact_kernel -- one thread per haystack walks the DP column by column with the exact (cost, length)
//    rules.  DP rows live in shared memory laid out [array][row][thread] (bank = thread, conflict-free).
//  * (lev_bitpar.cu) search_filter -- bit-parallel pre-filter for unit costs that flags the haystacks containing
//    at least one end position with cost <= k, so the exact kernel only runs on those.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "ta_common.cuh"
#include "search_emit.hpp"  // Hit, sort_hits, emit_matches: the host phase (plain C++, also compiled by the CPU tests)

// implemented in lev_bitpar.cu: flags[i] = 1 iff haystack i has an end position with unit-cost distance <= k
int ta_launch_search_filter(ta_ctx *ctx, const uint8_t *needle_dev, uint32_t needle_len, const uint8_t *hay,
                            const uint64_t *hay_off, size_t n, uint64_t max_hay, uint32_t k, bool transpose,
                            uint32_t *idx_out, uint32_t *counter, ta_filter_out *fo, cudaStream_t st);

namespace {

struct SearchArgs {
    const uint8_t *needle;  // device copy
    const uint8_t *hay;
    const uint64_t *hay_off;
    const uint32_t *idx;  // optional: work item w -> haystack idx[w] (or a segment code, see segs)
    uint32_t segs;        // 0: idx holds haystack indices.  > 0: idx holds haystack * segs + segment codes and the
                          // item covers only that TA_SEARCH_SUB-byte segment after a warm-up of `warm` bytes
    uint32_t gran;        // bytes of end positions per segment code
    const uint32_t *alt_flag;  // if set on the device: the list holds alt_segs / alt_gran codes (the fallback filter ran)
    uint32_t alt_segs, alt_gran;
    uint32_t warm;
    uint32_t split;       // segment items are cut into `split` parts of TA_SEARCH_SUB / split end positions, a warp each
    size_t n;
    const uint32_t *n_dev;  // optional: the number of work items lives on the device (written by the pre-filter)
    uint32_t needle_len;
    uint32_t k;
    uint32_t mism, gap, sgap, tcost;
    int anchored;
    Hit *hits;
    unsigned long long *hit_count;
    unsigned long long hit_cap;
    uint32_t *rows_ws;  // search_exact_kernel<.., GLOBAL = true>: DP rows of every block, (4 | 6) * (N + 1) * T words each
};

// GLOBAL = false: DP rows and the needle in shared memory (needles up to ~450 bytes).  GLOBAL = true: the rows live in a
// global-memory workspace with the same [array][row][thread] layout (coalesced across the threads of a block) and the
// needle is read through the read-only cache, for needles of any length the header promises (TA_MAX_STRING_LEN);
// blocks are persistent there because the workspace bounds how many can run.
template <bool TRANS, bool GLOBAL>
__global__ void __launch_bounds__(64) search_exact_kernel(const SearchArgs args) {
    extern __shared__ uint32_t sm_rows[];
    const uint32_t N = args.needle_len;
    const uint32_t rows = N + 1;
    const uint32_t T = blockDim.x;
    const uint32_t tid = threadIdx.x;
    uint32_t *const sm = GLOBAL ? args.rows_ws + (size_t)blockIdx.x * (TRANS ? 6 : 4) * rows * T : sm_rows;
    // array a, row j -> sm[(a * rows + j) * T + tid]
    auto at = [&](uint32_t a, uint32_t j) -> uint32_t & { return sm[((size_t)a * rows + j) * T + tid]; };
    enum { CUR_DP = 0, CUR_LEN = 1, NGAP = 2, NGAP_LEN = 3, PREV_DP = 4, PREV_LEN = 5 };
    const uint8_t *sneedle = args.needle;
    if (!GLOBAL) {
        uint8_t *sn = (uint8_t *)(sm_rows + (size_t)(TRANS ? 6 : 4) * rows * T);
        for (uint32_t q = tid; q < N; q += T) sn[q] = args.needle[q];
        __syncthreads();
        sneedle = sn;
    }

  for (size_t w = (size_t)blockIdx.x * T + tid; w < args.n; w += (size_t)gridDim.x * T) {
    const uint32_t hidx = args.idx ? args.idx[w] : (uint32_t)w;
    const uint64_t h0 = args.hay_off[hidx], h1 = args.hay_off[hidx + 1];
    const uint8_t *hay = args.hay + h0;
    const uint64_t H = h1 - h0;
    const uint32_t mism = args.mism, gap = args.gap, sgap = args.sgap, tcost = args.tcost, k = args.k;
    const uint32_t open = sgap + gap;
    const bool anchored = args.anchored != 0;

    // src/levenshtein.rs:1650-1661
    uint64_t iter_len = H;
    if (anchored) {
        const uint64_t lim = (uint64_t)N + (uint64_t)((k > sgap ? k - sgap : 0u) / gap);
        iter_len = H < lim ? H : lim;
    }

    // column 0 (src/levenshtein.rs:1689-1691 and the initial vectors :1663-1672)
    for (uint32_t j = 0; j <= N; j++) {
        at(CUR_DP, j) = j * gap + (j ? sgap : 0u);
        at(CUR_LEN, j) = 0;
        at(NGAP, j) = TA_INF;  // u32::MAX in the reference: any opened gap is cheaper
        at(NGAP_LEN, j) = 0;
        if (TRANS) {
            at(PREV_DP, j) = 0;
            at(PREV_LEN, j) = 0;
        }
    }

    uint32_t hc_prev = 0;
    for (uint64_t x = 1; x <= iter_len; x++) {  // column x consumes haystack[x-1]
        const uint32_t hc = __ldg(hay + (x - 1));
        const uint32_t row0 = anchored ? (uint32_t)x * gap + sgap : 0u;  // :1710-1721
        // (x-1, j-1) values before they are overwritten
        uint32_t diag_dp = at(CUR_DP, 0), diag_len = 0;
        // (x-2, j-2) pipeline for transpositions
        uint32_t t1_dp = 0, t1_len = 0, t2_dp = 0, t2_len = 0;
        if (TRANS) {
            t1_dp = at(PREV_DP, 0);  // dp(x-2, 0)
            t1_len = 0;
            at(PREV_DP, 0) = diag_dp;  // becomes dp(x-1, 0) for the next column
        }
        at(CUR_DP, 0) = row0;
        // (x, j-1) running values
        uint32_t left_dp = row0, left_len = 0, hgap = TA_INF, hgap_len = 0;
        uint32_t nprev = 0;  // needle[j-2]

        for (uint32_t j = 1; j <= N; j++) {
            const uint32_t nc = sneedle[j - 1];
            const uint32_t up_dp = at(CUR_DP, j), up_len = at(CUR_LEN, j);  // (x-1, j)
            const uint32_t sub = diag_dp + (nc != hc ? mism : 0u);
            const uint32_t sub_len = diag_len + 1;

            // needle gap: consume a haystack byte (:1726-1737)
            uint32_t ng = at(NGAP, j), ngl = at(NGAP_LEN, j);
            {
                const uint32_t new_gap = up_dp + open, cont_gap = ng + gap;
                if (new_gap < cont_gap) {
                    ng = new_gap;
                    ngl = up_len + 1;
                } else if (new_gap > cont_gap) {
                    ng = cont_gap;
                    ngl = ngl + 1;
                } else {
                    ng = cont_gap;
                    ngl = max(up_len, ngl) + 1;
                }
                ng = min(ng, TA_INF);
            }
            at(NGAP, j) = ng;
            at(NGAP_LEN, j) = ngl;

            // haystack gap: consume a needle byte (:1739-1750)
            {
                const uint32_t new_gap = left_dp + open, cont_gap = hgap + gap;
                if (new_gap < cont_gap) {
                    hgap = new_gap;
                    hgap_len = left_len;
                } else if (new_gap > cont_gap) {
                    hgap = cont_gap;
                    // hgap_len unchanged: haystack_gap_length[j] = haystack_gap_length[j-1]
                } else {
                    hgap = cont_gap;
                    hgap_len = max(left_len, hgap_len);
                }
                hgap = min(hgap, TA_INF);
            }

            uint32_t dp = ng, len = ngl;  // :1752-1753
            if (hgap < dp || (hgap == dp && left_len > len)) {  // :1755-1760 (compares length2[j-1])
                dp = hgap;
                len = hgap_len;
            }
            if (sub < dp || (sub == dp && sub_len > len)) {  // :1762-1765
                dp = sub;
                len = sub_len;
            }
            if (TRANS) {
                // (x-2, j-2): value read two rows ago
                const uint32_t p_dp = at(PREV_DP, j), p_len = at(PREV_LEN, j);  // dp(x-2, j), len(x-2, j)
                if (x > 1 && j > 1 && nc == hc_prev && nprev == hc) {           // :1767-1779
                    const uint32_t tr = t2_dp + tcost;
                    if (tr <= dp) {
                        dp = tr;
                        len = t2_len + 2;
                    }
                }
                t2_dp = t1_dp;
                t2_len = t1_len;
                t1_dp = p_dp;
                t1_len = p_len;
                at(PREV_DP, j) = up_dp;  // (x-1, j) becomes (x-2, j) two columns later
                at(PREV_LEN, j) = up_len;
            }
            at(CUR_DP, j) = dp;
            at(CUR_LEN, j) = len;
            diag_dp = up_dp;
            diag_len = up_len;
            left_dp = dp;
            left_len = len;
            nprev = nc;
        }
        hc_prev = hc;

        if (left_dp <= k) {  // :1792-1806 with the All threshold; Best is a host-side filter of this list
            const unsigned long long slot = atomicAdd(args.hit_count, 1ull);
            if (slot < args.hit_cap) {
                Hit h;
                h.hay = hidx;
                h.cost = left_dp;
                h.end = (uint32_t)x;
                h.len = (uint32_t)left_len;
                args.hits[slot] = h;
            }
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// search_wave_kernel: the same exact (cost, length) recurrence, one WARP per haystack.  Lane t owns needle rows
// t*C+1 .. t*C+C and is skewed t columns behind lane 0 (lane t handles haystack column x = s - t at step s), so the
// (x, j-1) neighbour state it needs was produced by lane t-1 one step earlier and arrives by __shfl_up_sync; the
// (x-1, j-1) and (x-2, j-2) values are the ones received one and two steps before.  Used when few haystacks survive
// the pre-filter (a thread-per-haystack walk would leave the GPU idle) and for needles of up to 32*8 rows.
template <int C, bool TRANS>
__global__ void __launch_bounds__(128) search_wave_kernel(const SearchArgs args) {
    const unsigned full = 0xffffffffu;
    const int t = threadIdx.x & 31;
    // persistent warps: the item count may only exist on the device (pre-filter output), so the grid is sized from an
    // upper bound and every warp strides over the items
    // A warp's item is ONE serial chain of warm + end positions + 31 steps (~75 dependent instructions each), and after
    // a pre-filter there are far fewer items than the GPU has warps -- so a flagged sub-segment is cut into `split`
    // parts, each restarted on its own: more total work, a shorter chain (37 + 32 + 31 steps instead of 37 + 128 + 31).
    uint32_t segs = args.segs, gran = args.gran;
    if (segs && args.alt_flag && *args.alt_flag) segs = args.alt_segs, gran = args.alt_gran;
    const uint32_t split = segs && gran >= (uint32_t)TA_SEARCH_SUB ? args.split : 1u;
    const size_t n_items = (args.n_dev ? (size_t)*args.n_dev : args.n) * split;
    const size_t n_warps = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_items; w += n_warps) {
    const uint32_t part = (uint32_t)(w % split);
    const uint32_t code = args.idx ? args.idx[w / split] : (uint32_t)w;
    const uint32_t hidx = segs ? code / segs : code;
    const uint64_t h0 = args.hay_off[hidx], h1 = args.hay_off[hidx + 1];
    const uint8_t *hay = args.hay + h0;
    uint64_t H = h1 - h0;
    const uint32_t N = args.needle_len;
    const uint32_t mism = args.mism, gap = args.gap, sgap = args.sgap, tcost = args.tcost, k = args.k;
    const uint32_t open = sgap + gap;
    const bool anchored = args.anchored != 0;
    // Segment mode (unanchored only): only cells of cost <= k are reported, a cell of cost c is decided by
    // predecessors of cost <= c (costs are non-negative), and a path of cost c to needle row j consumes at most
    // j + c/gap haystack bytes -- so every candidate that wins or ties in the reference's length tie-breaks lies
    // within N + k/gap bytes of the end position, and the candidates a restarted DP sees too expensively lose in both
    // runs: a fresh start `warm` = N + k/gap + 2 bytes before the segment reproduces the reference's (cost, length)
    // for every reported end position inside it (tests/test_search_restart_model.py pins the margin on the CPU).
    uint32_t col0 = 0, emit_from = 0;  // columns are numbered from col0; hits are reported for x + col0 > emit_from
    if (segs) {
        const uint32_t seg = code % segs, span = gran / split;
        emit_from = seg * gran + part * span;
        if (emit_from >= H) continue;  // this part lies past the haystack's end
        const uint64_t seg_end = (uint64_t)emit_from + span < H ? (uint64_t)emit_from + span : H;
        col0 = emit_from > args.warm ? emit_from - args.warm : 0;
        hay += col0;
        H = seg_end - col0;
    }
    uint32_t iter_len = (uint32_t)H;  // src/levenshtein.rs:1650-1661 (H <= 0xFFFFFF00: checked by the host)
    if (anchored) {
        const uint64_t lim = (uint64_t)N + (uint64_t)((k > sgap ? k - sgap : 0u) / gap);
        iter_len = H < lim ? (uint32_t)H : (uint32_t)lim;
    }

    // rows owned by this lane: j = t*C + c + 1
    uint32_t nc[C], dp1[C], len1[C], ng[C], ngl[C], dp2[C], len2[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const uint32_t j = (uint32_t)(t * C + c + 1);
        nc[c] = j <= N ? (uint32_t)args.needle[j - 1] : 0x100u;  // rows past the needle never match
        dp1[c] = j * gap + sgap;                                   // column 0 (:1689-1691)
        len1[c] = 0;
        ng[c] = TA_INF;
        ngl[c] = 0;
        dp2[c] = 0;
        len2[c] = 0;
    }
    const uint32_t nprev0 = (t * C >= 1 && (uint32_t)(t * C) <= N) ? (uint32_t)args.needle[t * C - 1] : 0x100u;

    // state of the row just above this lane's rows (row t*C): received values for columns x, x-1, x-2
    uint32_t Ldp1 = (uint32_t)(t * C) * gap + (t ? sgap : 0u), Llen1 = 0;  // column x-1 (starts as column 0)
    uint32_t Ldp2 = 0, Llen2 = 0;                                            // column x-2
    uint32_t T0dp1 = 0, T0len1 = 0;  // (x-2, row t*C-1) once shifted; pipeline of the value forwarded for TRANS
    // what this lane offers to lane t+1: its last row at the column it just finished
    uint32_t out_dp = dp1[C - 1], out_len = 0, out_hg = TA_INF, out_hgl = 0;
    uint32_t out_tdp = 0, out_tlen = 0;  // (col-1 of the finished column, row (t+1)*C - 1) for lane t+1's TRANS
    uint32_t hc_prev = 0x200u;

    const int last_t = (int)((N - 1) / C), last_c = (int)((N - 1) % C);
    const uint32_t steps = iter_len + 31;  // 32-bit counters: the step is one serial chain, every instruction counts
    // haystack bytes go through registers: chunk q = bytes [32q, 32q + 32), one per lane, fetched a chunk ahead.  At step
    // s the lanes read bytes s-1-t, which lie in chunk (s-1)/32 or the one before -- two shuffles instead of a dependent
    // one-byte global load in every step of the serial chain (that load was 80 % of the step: 48 -> ~10 us per launch).
    auto ld_chunk = [&](const uint32_t q) {
        const uint32_t i = q * 32 + (uint32_t)t;
        return i < iter_len ? (uint32_t)__ldg(hay + i) : 0u;
    };
    uint32_t hb_prev = 0, hb_cur = ld_chunk(0), hb_next = ld_chunk(1);
    for (uint32_t s = 1; s <= steps; s++) {
        if (s > 1 && ((s - 1) & 31) == 0) {
            hb_prev = hb_cur;
            hb_cur = hb_next;
            hb_next = ld_chunk(((s - 1) >> 5) + 1);
        }
        // byte index of column x is s - 1 - t: in the current chunk iff t <= (s - 1) mod 32
        const int src_lane = (int)((s - 1 - (uint32_t)t) & 31u);
        const uint32_t hb_a = __shfl_sync(full, hb_cur, src_lane), hb_b = __shfl_sync(full, hb_prev, src_lane);
        const uint32_t hc = (uint32_t)t <= ((s - 1) & 31u) ? hb_a : hb_b;
        // neighbour state for column x = s - t, produced by lane t-1 at step s-1
        uint32_t Ldp = __shfl_up_sync(full, out_dp, 1);
        uint32_t Llen = __shfl_up_sync(full, out_len, 1);
        uint32_t Lhg = __shfl_up_sync(full, out_hg, 1);
        uint32_t Lhgl = __shfl_up_sync(full, out_hgl, 1);
        uint32_t Tdp = 0, Tlen = 0;
        if (TRANS) {
            Tdp = __shfl_up_sync(full, out_tdp, 1);
            Tlen = __shfl_up_sync(full, out_tlen, 1);
        }
        const uint32_t x = s - (uint32_t)t;  // meaningful when s > t
        const bool active = s > (uint32_t)t && x <= iter_len;
        if (t == 0) {  // row 0 (:1710-1721)
            Ldp = anchored ? (uint32_t)x * gap + sgap : 0u;
            Llen = 0;
            Lhg = TA_INF;
            Lhgl = 0;
        }
        if (active) {
            // running (x, j-1) values, (x-1, j-1) diagonal, (x-2, j-2) for transpositions
            uint32_t left_dp = Ldp, left_len = Llen, hgap = Lhg, hgap_len = Lhgl;
            uint32_t diag_dp = Ldp1, diag_len = Llen1;
            uint32_t tr_dp = T0dp1, tr_len = T0len1;      // (x-2, row t*C - 1): for c == 0
            uint32_t tr_dp_next = Ldp2, tr_len_next = Llen2;  // (x-2, row t*C): for c == 1
            uint32_t nprev = nprev0;
            uint32_t new_out_tdp = 0, new_out_tlen = 0;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const uint32_t up_dp = dp1[c], up_len = len1[c];
                const uint32_t sub = diag_dp + (nc[c] != hc ? mism : 0u);
                const uint32_t sub_len = diag_len + 1;
                {  // needle gap (:1726-1737)
                    const uint32_t new_gap = up_dp + open, cont_gap = ng[c] + gap;
                    if (new_gap < cont_gap) {
                        ng[c] = new_gap;
                        ngl[c] = up_len + 1;
                    } else if (new_gap > cont_gap) {
                        ng[c] = cont_gap;
                        ngl[c] = ngl[c] + 1;
                    } else {
                        ng[c] = cont_gap;
                        ngl[c] = max(up_len, ngl[c]) + 1;
                    }
                    ng[c] = min(ng[c], TA_INF);
                }
                {  // haystack gap (:1739-1750)
                    const uint32_t new_gap = left_dp + open, cont_gap = hgap + gap;
                    if (new_gap < cont_gap) {
                        hgap = new_gap;
                        hgap_len = left_len;
                    } else if (new_gap > cont_gap) {
                        hgap = cont_gap;
                    } else {
                        hgap = cont_gap;
                        hgap_len = max(left_len, hgap_len);
                    }
                    hgap = min(hgap, TA_INF);
                }
                uint32_t dp = ng[c], len = ngl[c];
                if (hgap < dp || (hgap == dp && left_len > len)) {  // :1755-1760
                    dp = hgap;
                    len = hgap_len;
                }
                if (sub < dp || (sub == dp && sub_len > len)) {  // :1762-1765
                    dp = sub;
                    len = sub_len;
                }
                if (TRANS) {
                    const uint32_t j = (uint32_t)(t * C + c + 1);
                    if (x > 1 && j > 1 && nc[c] == hc_prev && nprev == hc) {  // :1767-1779
                        const uint32_t tr = tr_dp + tcost;
                        if (tr <= dp) {
                            dp = tr;
                            len = tr_len + 2;
                        }
                    }
                    // next cell's (x-2, j-2) is this cell's row - 1 at column x-2
                    tr_dp = tr_dp_next;
                    tr_len = tr_len_next;
                    tr_dp_next = dp2[c];
                    tr_len_next = len2[c];
                    if (c == C - 2) {  // row (t+1)*C - 1 at column x-1: what lane t+1's first cell will need
                        new_out_tdp = up_dp;
                        new_out_tlen = up_len;
                    }
                    dp2[c] = up_dp;
                    len2[c] = up_len;
                }
                dp1[c] = dp;
                len1[c] = len;
                diag_dp = up_dp;
                diag_len = up_len;
                left_dp = dp;
                left_len = len;
                nprev = nc[c];
                if (c == last_c && t == last_t && dp <= k && x + col0 > emit_from) {
                    const unsigned long long slot = atomicAdd(args.hit_count, 1ull);  // :1792-1806 (All threshold)
                    if (slot < args.hit_cap) {
                        Hit h;
                        h.hay = hidx;
                        h.cost = dp;
                        h.end = x + col0;
                        h.len = (uint32_t)len;
                        args.hits[slot] = h;
                    }
                }
            }
            if (TRANS) {
                if (C == 1) {  // row t at column x-1 is the row above this lane's single row: forward L(x-1)
                    new_out_tdp = Ldp1;
                    new_out_tlen = Llen1;
                }
                out_tdp = new_out_tdp;
                out_tlen = new_out_tlen;
                // shift the received-history pipelines
                T0dp1 = Tdp;  // received now: (x-1, row t*C-1); next step it is (x-2, ...) relative to column x+1
                T0len1 = Tlen;
                Ldp2 = Ldp1;
                Llen2 = Llen1;
                hc_prev = hc;
            }
            Ldp1 = Ldp;
            Llen1 = Llen;
            out_dp = left_dp;
            out_len = left_len;
            out_hg = hgap;
            out_hgl = hgap_len;
        }
    }
    }  // work items
}

size_t exact_smem_bytes(uint32_t needle_len, bool trans, int threads) {
    return (size_t)(trans ? 6 : 4) * (needle_len + 1) * threads * sizeof(uint32_t) + ((needle_len + 15) & ~15u);
}

}  // namespace

// Device phase: optional bit-parallel pre-filter, then the exact kernel on the surviving segments / haystacks; returns
// every end position with cost <= k as a Hit (unordered).  Caller holds ctx->mu and has set the device.
// With the pre-filter the whole phase is queued without a host round trip: the filter appends the codes of the
// segments that contain a match end to a device list, the (persistent) wave kernel reads the list's length from the
// device, and one synchronisation at the end brings back both counters together with the first SPEC_HITS hits.
static int search_device(ta_ctx *ctx, cudaStream_t st, const uint8_t *d_needle, size_t needle_len,
                         const uint8_t *d_hay, const uint64_t *d_off, size_t n, uint64_t max_hay, uint32_t k,
                         ta_costs costs, int anchored, std::vector<Hit> &hits) {
    int rc;
    if (max_hay > 0xFFFFFF00ull) return TA_ERR_TOO_LARGE;  // hit records and the kernels' column counters are 32-bit
    constexpr size_t SPEC_HITS = 16384;  // hits copied back speculatively with the counters (384 KB, pinned)
    const bool unit = costs.mismatch == 1 && costs.gap == 1 && costs.start_gap == 0 && costs.transpose <= 1;
    const uint32_t *d_idx = nullptr;
    const uint32_t *d_work_n = nullptr;  // device-side item count (pre-filter path)
    size_t work_n = n;                   // items, or their upper bound while the count is still on the device
    uint32_t segs = 0;                   // > 0: work items are flagged haystack segments (pre-filter ran)
    ta_filter_out fo;
    static const bool no_filter = getenv("TA_NO_SEARCH_FILTER") != nullptr;  // testing: exact kernel on everything
    uint32_t *counter = ctx->d_flags + 2;
    unsigned long long *d_count = (unsigned long long *)(ctx->d_flags + 4);
    // item counter, pad, 64-bit hit counter, and the q-gram filter's queue length and gave-up flag (d_flags + 6, + 7)
    TA_CUDA(ctx, cudaMemsetAsync(counter, 0, 6 * sizeof(uint32_t), st));
    // The pre-filters work with unit costs.  For any other cost model they run with the number of edit OPERATIONS a
    // match can contain: an alignment of weighted cost <= k has at most ku = k / min(mismatch, gap, transpose) operations,
    // so its end position also ends a unit-cost alignment of cost <= ku -- the flagged sub-segments are a superset of
    // the weighted matches, and the exact kernel (which computes the real costs) decides.
    uint32_t cmin = costs.mismatch < costs.gap ? costs.mismatch : costs.gap;
    if (costs.transpose && costs.transpose < cmin) cmin = costs.transpose;
    const uint32_t ku = unit ? k : k / cmin;
    if (!no_filter && needle_len <= 64 && !anchored && ku < needle_len) {
        const uint64_t nseg = max_hay ? (max_hay + TA_SEARCH_SUB - 1) / TA_SEARCH_SUB : 1;
        if ((uint64_t)n * nseg <= 0xFFFFFFF0ull && nseg <= 65535) {
            // list capacity: every TA_SEARCH_SUB granule of every haystack (the scanning filters' worst case), or what the
            // q-gram resolve kernel can append (<= 8 distinct 16-byte granules per queue entry)
            const uint64_t qgram_items = 8ull * std::min<uint64_t>((uint64_t)n * max_hay / TA_QGRAM_BYTES_PER_ENTRY + 4096, 1ull << 24);
            const size_t list_cap = (size_t)std::max<uint64_t>((uint64_t)n * nseg, qgram_items);
            if ((rc = ta_dev_reserve(ctx, ctx->d_work[0], list_cap * sizeof(uint32_t))) != TA_OK) return rc;
            rc = ta_launch_search_filter(ctx, d_needle, (uint32_t)needle_len, d_hay, d_off, n, max_hay, ku,
                                         costs.transpose != 0, (uint32_t *)ctx->d_work[0].p, counter, &fo, st);
            segs = fo.segs;
            if (rc == TA_OK) {
                work_n = list_cap;
                d_work_n = counter;
                d_idx = (const uint32_t *)ctx->d_work[0].p;
            } else if (rc == TA_ERR_TOO_LARGE) {
                segs = 0;
            } else {
                return rc;
            }
        }
    }
    if (work_n == 0) return TA_OK;

    const bool trans = costs.transpose != 0;
    // kernel choice: warp-per-haystack wavefront when the thread-per-haystack walk could not fill the GPU (or its
    // shared-memory rows would not fit); TA_SEARCH_KERNEL=thread|wave forces one (testing)
    static const char *force = getenv("TA_SEARCH_KERNEL");
    int threads = 64;
    size_t smem = exact_smem_bytes((uint32_t)needle_len, trans, threads);
    if (smem > (size_t)ctx->smem_optin) {
        threads = 32;
        smem = exact_smem_bytes((uint32_t)needle_len, trans, threads);
    }
    const bool thread_ok = smem <= (size_t)ctx->smem_optin;
    const bool wave_ok = needle_len <= 256;
    bool use_wave = wave_ok && (!thread_ok || work_n < (size_t)ctx->sm_count * 256);
    if (force && force[0] == 't' && thread_ok) use_wave = false;
    if (force && force[0] == 'w' && wave_ok) use_wave = true;
    if (segs) use_wave = true;  // segment work items are only understood by the wave kernel (needle <= 64 here)
    // needles that fit neither kernel (> 256 bytes and DP rows beyond shared memory; the reference takes any needle,
    // src/levenshtein.rs:2034-2151): thread per haystack with the DP rows in a global-memory workspace
    const bool use_global = (!use_wave && !thread_ok) || (force && force[0] == 'g' && !segs);
    size_t global_blocks = 0;
    if (use_global) {
        use_wave = false;
        threads = 32;
        smem = 0;
        const size_t per_block = (size_t)(trans ? 6 : 4) * (needle_len + 1) * threads * sizeof(uint32_t);
        const size_t budget = (size_t)2 << 30;
        global_blocks = std::max<size_t>(1, std::min<size_t>({budget / per_block, (work_n + threads - 1) / threads,
                                                              (size_t)ctx->sm_count * 16}));
        if ((rc = ta_dev_reserve(ctx, ctx->d_work[3], global_blocks * per_block)) != TA_OK) return rc;
    }
    void (*kern)(const SearchArgs) = nullptr;
    if (use_wave) {
        const int C = needle_len <= 32 ? 1 : needle_len <= 64 ? 2 : needle_len <= 128 ? 4 : 8;
        if (C == 1) kern = trans ? search_wave_kernel<1, true> : search_wave_kernel<1, false>;
        if (C == 2) kern = trans ? search_wave_kernel<2, true> : search_wave_kernel<2, false>;
        if (C == 4) kern = trans ? search_wave_kernel<4, true> : search_wave_kernel<4, false>;
        if (C == 8) kern = trans ? search_wave_kernel<8, true> : search_wave_kernel<8, false>;
        threads = 128;
        smem = 0;
    } else if (use_global) {
        kern = trans ? search_exact_kernel<true, true> : search_exact_kernel<false, true>;
    } else {
        kern = trans ? search_exact_kernel<true, false> : search_exact_kernel<false, false>;
        TA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if ((rc = ta_pin_reserve(ctx, ctx->h_pin[3], SPEC_HITS * sizeof(Hit))) != TA_OK) return rc;
    // hit buffer: sized from the item count when the host knows it, else a generous default; an overflow is detected
    // from the final counter and the exact kernel re-run once with the exact size
    unsigned long long cap = d_work_n ? std::max<unsigned long long>(1ull << 18, n / 4)
                                      : std::max<unsigned long long>(4096, work_n * 8);
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = ta_dev_reserve(ctx, ctx->d_work[1], cap * sizeof(Hit))) != TA_OK) return rc;
        if (attempt) TA_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
        SearchArgs sa;
        sa.needle = d_needle, sa.hay = d_hay, sa.hay_off = d_off, sa.idx = d_idx, sa.n = work_n, sa.n_dev = d_work_n;
        sa.segs = segs, sa.warm = (uint32_t)needle_len + k / costs.gap + 2u;
        sa.gran = fo.gran, sa.alt_flag = fo.alt_flag, sa.alt_segs = fo.alt_segs, sa.alt_gran = fo.alt_gran;
        static const int env_split = getenv("TA_WAVE_SPLIT") ? atoi(getenv("TA_WAVE_SPLIT")) : 0;  // 1, 2, 4, 8 (testing)
        sa.split = (env_split == 1 || env_split == 2 || env_split == 4 || env_split == 8) ? (uint32_t)env_split : 2u;
        sa.needle_len = (uint32_t)needle_len, sa.k = k;
        sa.mism = costs.mismatch, sa.gap = costs.gap, sa.sgap = costs.start_gap, sa.tcost = costs.transpose;
        sa.anchored = anchored, sa.hits = (Hit *)ctx->d_work[1].p, sa.hit_count = d_count, sa.hit_cap = cap;
        sa.rows_ws = use_global ? (uint32_t *)ctx->d_work[3].p : nullptr;
        const size_t per_block = use_wave ? (size_t)threads / 32 : (size_t)threads;
        size_t blocks = (work_n * (segs ? sa.split : 1u) + per_block - 1) / per_block;
        if (use_wave) blocks = std::min<size_t>(blocks, (size_t)ctx->sm_count * 16);  // persistent warps
        if (use_global) blocks = global_blocks;
        kern<<<(unsigned)blocks, threads, smem, st>>>(sa);
        ctx->launches++;
        TA_CUDA(ctx, cudaGetLastError());
        const size_t spec = (size_t)std::min<unsigned long long>(cap, SPEC_HITS);
        TA_CUDA(ctx, cudaMemcpyAsync(ctx->h_flags + 2, counter, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        TA_CUDA(ctx, cudaMemcpyAsync(ctx->h_pin[3].p, ctx->d_work[1].p, spec * sizeof(Hit), cudaMemcpyDeviceToHost, st));
        TA_CUDA(ctx, cudaStreamSynchronize(st));
        unsigned long long got;
        memcpy(&got, ctx->h_flags + 4, sizeof got);
        if (got <= cap) {
            hits.resize((size_t)got);
            const size_t head = (size_t)std::min<unsigned long long>(got, spec);
            if (head) memcpy(hits.data(), ctx->h_pin[3].p, head * sizeof(Hit));
            if (got > head) {
                TA_CUDA(ctx, cudaMemcpyAsync(hits.data() + head, (const Hit *)ctx->d_work[1].p + head,
                                             (size_t)(got - head) * sizeof(Hit), cudaMemcpyDeviceToHost, st));
                TA_CUDA(ctx, cudaStreamSynchronize(st));
            }
            return TA_OK;
        }
        cap = got;  // exact size known now: rerun once
    }
    return TA_ERR_TOO_LARGE;
}

static int export_matches(const std::vector<ta_match> &result, uint64_t *moff, ta_match **out_matches,
                          uint64_t **out_match_off) {
    ta_match *m = (ta_match *)ta_out_alloc((result.size() ? result.size() : 1) * sizeof(ta_match));
    if (!m) {
        ta_free(moff);
        return TA_ERR_NOMEM;
    }
    if (!result.empty()) memcpy(m, result.data(), result.size() * sizeof(ta_match));
    *out_matches = m;
    *out_match_off = moff;
    return TA_OK;
}

// Device + emit phases for the haystacks hay_off[0 .. n] on ONE device: uploads the haystacks (and the needle, unless
// `needle` is null = it is already in ctx->d_b[0], put there by the multi-device broadcast), runs the kernels, and
// fills moff[0 .. n] (moff[0] = 0 on entry) and `result` with this range's matches.  The caller holds ctx->mu.
static int search_one_device(ta_ctx *ctx, const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                             const uint64_t *hay_off, size_t n, uint32_t k, bool best, ta_costs costs, int anchored,
                             uint64_t *moff, std::vector<ta_match> &result) {
    uint64_t max_hay = 0;
    for (size_t i = 0; i < n; i++) max_hay = std::max(max_hay, hay_off[i + 1] - hay_off[i]);
    const uint64_t total_hay = hay_off[n] - hay_off[0];
    std::vector<Hit> hits;
    auto run = [&]() -> int {
        TA_CUDA(ctx, cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        int rc;
        const uint64_t lo = hay_off[0];
        const size_t skew = (size_t)(lo & 15);
        if ((rc = ta_dev_reserve(ctx, ctx->d_a[0], skew + total_hay + 64)) != TA_OK) return rc;
        if ((rc = ta_dev_reserve(ctx, ctx->d_aoff[0], (n + 1) * sizeof(uint64_t))) != TA_OK) return rc;
        if (needle && (rc = ta_dev_reserve(ctx, ctx->d_b[0], needle_len + 64)) != TA_OK) return rc;
        if (total_hay)
            TA_CUDA(ctx, cudaMemcpyAsync((uint8_t *)ctx->d_a[0].p + skew, hay + lo, total_hay, cudaMemcpyHostToDevice, st));
        TA_CUDA(ctx, cudaMemcpyAsync(ctx->d_aoff[0].p, hay_off, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        if (needle) TA_CUDA(ctx, cudaMemcpyAsync(ctx->d_b[0].p, needle, needle_len, cudaMemcpyHostToDevice, st));
        return search_device(ctx, st, (const uint8_t *)ctx->d_b[0].p, needle_len,
                             (const uint8_t *)ctx->d_a[0].p + skew - lo, (const uint64_t *)ctx->d_aoff[0].p, n, max_hay, k,
                             costs, anchored, hits);
    };
    const int rc = run();
    if (rc != TA_OK) {
        cudaStreamSynchronize(ctx->stream);
        return rc;
    }
    emit_matches(n, needle_len, k, best, costs, hits, moff, result);
    return TA_OK;
}

extern "C" int ta_levenshtein_search_batch(ta_ctx *ctx, const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                                           const uint64_t *hay_off, size_t n, uint32_t k, int search_type,
                                           ta_costs costs, int anchored, ta_match **out_matches,
                                           uint64_t **out_match_off) {
    if (!ctx || !out_matches || !out_match_off) return TA_ERR_BAD_ARG;
    *out_matches = nullptr;
    *out_match_off = nullptr;
    if (search_type != TA_SEARCH_ALL && search_type != TA_SEARCH_BEST) return TA_ERR_BAD_ARG;
    if (!ta_costs_valid(costs)) return TA_ERR_BAD_COSTS;
    if (n && !hay_off) return TA_ERR_BAD_ARG;
    if (needle_len && !needle) return TA_ERR_BAD_ARG;
    if (n > 0xFFFFFFF0ull || needle_len > TA_MAX_STRING_LEN) return TA_ERR_TOO_LARGE;
    const bool best = search_type == TA_SEARCH_BEST;

    uint64_t max_hay = 0, total_hay = 0;
    for (size_t i = 0; i < n; i++) {
        if (hay_off[i + 1] < hay_off[i]) return TA_ERR_BAD_ARG;
        max_hay = std::max(max_hay, hay_off[i + 1] - hay_off[i]);
    }
    if (n) total_hay = hay_off[n] - hay_off[0];
    if (total_hay && !hay) return TA_ERR_BAD_ARG;

    uint64_t *moff = (uint64_t *)ta_out_alloc((n + 1) * sizeof(uint64_t));
    if (moff) moff[0] = 0;
    if (!moff) return TA_ERR_NOMEM;
    std::vector<ta_match> result;

    if (needle_len == 0) {  // reference src/levenshtein.rs:1600-1644 -- no DP involved, pure bookkeeping
        for (size_t i = 0; i < n; i++) {
            if (anchored) {
                result.push_back(ta_match{0, 0, 0, 0});
                if (!best) {
                    const uint64_t H = hay_off[i + 1] - hay_off[i];
                    uint32_t cost = costs.start_gap;
                    for (uint64_t x = 0; x < H; x++) {
                        cost += costs.gap;
                        if (cost <= k)
                            result.push_back(ta_match{0, x + 1, cost, 0});
                        else
                            break;
                    }
                }
            }
            moff[i + 1] = result.size();
        }
        return export_matches(result, moff, out_matches, out_match_off);
    }
    if (!ta_costs_valid_search(costs)) {  // src/levenshtein.rs:1647
        ta_free(moff);
        return TA_ERR_BAD_COSTS;
    }
    if (n == 0) return export_matches(result, moff, out_matches, out_match_off);

    if (ctx->multi) {
        // one call, several GPUs: haystacks in contiguous ranges balanced by bytes, the needle broadcast from the first
        // device with NCCL, per-range match lists concatenated in range order
        std::lock_guard<std::mutex> lock(ctx->mu);
        const int parts = ta_multi_parts(ctx, total_hay, n);
        std::vector<size_t> bound;
        ta_multi_bounds(hay_off, nullptr, n, parts, bound);
        int rc = ta_multi_needle(ctx, needle, needle_len, parts);
        std::vector<ta_match *> ms(parts, nullptr);
        std::vector<uint64_t *> mo(parts, nullptr);
        if (rc == TA_OK)
            rc = ta_multi_run(ctx, parts, [&](int r) -> int {
                const size_t lo = bound[r], cnt = bound[r + 1] - bound[r];
                if (cnt == 0) return TA_OK;
                ta_ctx *sub = ta_multi_sub(ctx, r);
                uint64_t *lm = (uint64_t *)ta_out_alloc((cnt + 1) * sizeof(uint64_t));
                if (!lm) return TA_ERR_NOMEM;
                lm[0] = 0;
                std::vector<ta_match> res;
                int rr = search_one_device(sub, nullptr, needle_len, hay, hay_off + lo, cnt, k, best, costs, anchored, lm, res);
                if (rr != TA_OK) {
                    ta_free(lm);
                    return rr;
                }
                return export_matches(res, lm, &ms[r], &mo[r]);
            });
        ta_free(moff);
        if (rc != TA_OK) {
            for (int r = 0; r < parts; r++) ta_free(ms[r]), ta_free(mo[r]);
            return rc;
        }
        return ta_concat_lists<ta_match>(parts, bound, n, ms, mo, out_matches, out_match_off);
    }
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        const int rc = search_one_device(ctx, needle, needle_len, hay, hay_off, n, k, best, costs, anchored, moff, result);
        if (rc != TA_OK) {
            ta_free(moff);
            return rc;
        }
    }
    return export_matches(result, moff, out_matches, out_match_off);
}

extern "C" int ta_levenshtein_search_batch_dev(ta_ctx *ctx, const uint8_t *needle, size_t needle_len,
                                               const uint8_t *hay, const uint64_t *hay_off, size_t n,
                                               uint64_t max_hay_len, uint32_t k, int search_type, ta_costs costs,
                                               int anchored, ta_match **out_matches, uint64_t **out_match_off,
                                               void *stream) {
    if (!ctx || ctx->multi || !out_matches || !out_match_off) return TA_ERR_BAD_ARG;
    *out_matches = nullptr;
    *out_match_off = nullptr;
    if (search_type != TA_SEARCH_ALL && search_type != TA_SEARCH_BEST) return TA_ERR_BAD_ARG;
    if (!ta_costs_valid(costs)) return TA_ERR_BAD_COSTS;
    if (needle_len == 0 || !needle) return TA_ERR_BAD_ARG;  // the empty-needle bookkeeping needs host offsets
    if (!ta_costs_valid_search(costs)) return TA_ERR_BAD_COSTS;
    if (n && (!hay_off || !hay)) return TA_ERR_BAD_ARG;
    if (n > 0xFFFFFFF0ull || needle_len > TA_MAX_STRING_LEN) return TA_ERR_TOO_LARGE;
    uint64_t *moff = (uint64_t *)ta_out_alloc((n + 1) * sizeof(uint64_t));
    if (moff) moff[0] = 0;
    if (!moff) return TA_ERR_NOMEM;
    std::vector<ta_match> result;
    if (n == 0) return export_matches(result, moff, out_matches, out_match_off);
    std::vector<Hit> hits;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        auto run = [&]() -> int {
            TA_CUDA(ctx, cudaSetDevice(ctx->device));
            cudaStream_t st = (cudaStream_t)stream;
            int rc;
            if ((rc = ta_dev_reserve(ctx, ctx->d_b[0], needle_len + 64)) != TA_OK) return rc;
            TA_CUDA(ctx, cudaMemcpyAsync(ctx->d_b[0].p, needle, needle_len, cudaMemcpyHostToDevice, st));
            return search_device(ctx, st, (const uint8_t *)ctx->d_b[0].p, needle_len, hay, hay_off, n, max_hay_len, k,
                                 costs, anchored, hits);
        };
        static const bool trace = getenv("TA_TRACE_SEARCH") != nullptr;  // host-side phase times (profiling aid)
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = run();
        if (rc != TA_OK) {
            cudaStreamSynchronize((cudaStream_t)stream);
            ta_free(moff);
            return rc;
        }
        if (trace) {
            const auto t1 = std::chrono::steady_clock::now();
            emit_matches(n, needle_len, k, search_type == TA_SEARCH_BEST, costs, hits, moff, result);
            const auto t2 = std::chrono::steady_clock::now();
            const int rc2 = export_matches(result, moff, out_matches, out_match_off);
            const auto t3 = std::chrono::steady_clock::now();
            auto us = [](auto x, auto y) { return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(y - x).count() / 1e3; };
            fprintf(stderr, "[ta search] device phase %.1f us, emit %.1f us (%zu hits), export %.1f us (%zu matches)\n",
                    us(t0, t1), us(t1, t2), hits.size(), us(t2, t3), result.size());
            return rc2;
        }
    }
    emit_matches(n, needle_len, k, search_type == TA_SEARCH_BEST, costs, hits, moff, result);
    return export_matches(result, moff, out_matches, out_match_off);
}
