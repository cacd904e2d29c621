// lev_diag16.cu -- general EditCosts, narrow bands, large batches: one THREAD per pair, the band's anti-diagonals in
// registers as packed u16x2 cells, DPX min/add-min instructions (VIADDMNMX.U16x2, VIMNMX3.U16x2) on sm_100a.
//
// Contract: levenshtein_simd_k_with_opts / levenshtein_naive_k_with_opts (reference src/levenshtein.rs:714-827,
// 376-545) for every EditCosts {mismatch, gap, start_gap (affine), transpose}: out = d if d <= k else TA_NONE.  The
// reference picks the cell width from the input (u8 / u16 / u32 lanes, src/levenshtein.rs:767-790); this kernel is
// the u16 member of that family: the dispatcher (ta_launch_lev_band, lev_band.cu) uses it when max_k < 0x7F00, the
// band fits 32 diagonals and the batch is large enough to give every SM thousands of pairs; everything else goes to
// the lane-group kernel lev_band_kernel (u32 cells).
//
// Same recurrence and the same Ukkonen band [dlo, dlo + W) as lev_band_kernel (band_info in lev_band_common.cuh), but
// nothing is exchanged between lanes: a thread owns all 4 NR diagonals of its pair.  Cells of one anti-diagonal
// s = i + j have the same parity of (diagonal - dlo), so the band splits into an "even" array E (diagonals dlo + 2q)
// and an "odd" array O (dlo + 2q + 1) that are updated alternately; cell q of either array sits at row i0 - q, so the
// characters of `a` run backwards and those of `b` forwards along an anti-diagonal: two byte windows (NW registers
// each) that shift by one byte per step.  Per packed register (two cells) and step: byte compare of the windows
// (3 instructions per 4 cells), mask expansion (PRMT), substitution cost (LOP3), diag + cost clamped (VIADDMNMX), the
// left / upper neighbour pair (PRMT), min3 (VIMNMX3) and, with affine gaps, the two pre-minimised gap offers
// (VIADD + 2 VIADDMNMX).
//
// Why clamping is exact: all costs are non-negative, so a cell whose value exceeds max_k cannot lie on a path of
// cost <= max_k; clamping every cell to CAP = 0x7FFF > max_k therefore changes no answer that is reported, and
// keeps every intermediate below 2^16 (cells <= CAP, offers <= CAP + 510).  Cells outside the matrix start at CAP
// and stay there; D(0,0) = 0 is planted where the first anti-diagonal meets diagonal 0; the first row and column
// come out of the affine gap chains on their own (D(0,j) = start_gap + j gap is exactly the chain opened at (0,0)).
//
// Strings are streamed from global memory one aligned 32-bit word per four columns and string, re-aligned with a
// funnel shift (any CSR offset); words outside the string are never loaded.
#include <stdlib.h>

#include <algorithm>

#include "lev_band_common.cuh"
#include "lev_diag16_core.cuh"
#include "ta_common.cuh"

namespace {

template <int NR, bool AFFINE, bool TRANS>
__global__ void __launch_bounds__(128) lev_diag16_kernel(const BandArgs args) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= args.n) return;
    const size_t pair = args.idx ? (size_t)args.idx[w] : args.pair_base + w;
    const uint64_t a0 = args.a_off[pair], a1 = args.a_off[pair + 1];
    const uint64_t b0 = args.b_off[pair], b1 = args.b_off[pair + 1];
    args.out[pair] = diag16::pair<NR, AFFINE, TRANS>(args.a + a0, a1 - a0, args.b + b0, b1 - b0, args.k, args.mism,
                                                     args.gap, args.sgap, args.tcost);
}

template <int NR>
int launch_nr(ta_ctx *ctx, const BandArgs &args, bool affine, bool trans, cudaStream_t st) {
    const unsigned blocks = (unsigned)((args.n + 127) / 128);
    if (affine && trans)
        lev_diag16_kernel<NR, true, true><<<blocks, 128, 0, st>>>(args);
    else if (affine)
        lev_diag16_kernel<NR, true, false><<<blocks, 128, 0, st>>>(args);
    else if (trans)
        lev_diag16_kernel<NR, false, true><<<blocks, 128, 0, st>>>(args);
    else
        lev_diag16_kernel<NR, false, false><<<blocks, 128, 0, st>>>(args);
    ctx->launches++;
    TA_CUDA(ctx, cudaGetLastError());
    return TA_OK;
}

}  // namespace

// Does the thread-per-pair u16 kernel take this batch?  W = bound on the band width of any pair (ta_band_width_bound).
bool ta_diag16_can_handle(size_t n, uint32_t k, ta_costs c, uint32_t max_len, uint32_t W) {
    static const int force = getenv("TA_DIAG16") ? atoi(getenv("TA_DIAG16")) : -1;  // 0 = never, 1 = whenever valid
    if (force == 0) return false;
    if (W > 32) return false;
    // cells are clamped to 0x7FFF: every reportable distance must stay below it
    const uint64_t ub = (uint64_t)max_len * (c.mismatch > c.gap ? c.mismatch : c.gap) + c.start_gap;
    const uint64_t kk = k < ub ? k : ub;
    if (kk >= 0x7F00u) return false;
    if (force == 1) return true;
    return n >= 16384;  // a thread per pair: small batches are better off with a lane group per pair
}

int ta_launch_lev_diag16(ta_ctx *ctx, const BandArgs &args, ta_costs costs, uint32_t W, cudaStream_t st) {
    const bool affine = costs.start_gap != 0, trans = costs.transpose != 0;
    if (W <= 8) return launch_nr<2>(ctx, args, affine, trans, st);
    if (W <= 12) return launch_nr<3>(ctx, args, affine, trans, st);
    if (W <= 16) return launch_nr<4>(ctx, args, affine, trans, st);
    if (W <= 24) return launch_nr<6>(ctx, args, affine, trans, st);
    return launch_nr<8>(ctx, args, affine, trans, st);
}
