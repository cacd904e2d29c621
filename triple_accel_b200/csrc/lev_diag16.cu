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

// `defer` (may be null): pairs whose own band does not fit 4 NR diagonals are appended to defer[1 ..] (count in
// defer[0]) instead of being computed; `n_dev` (may be null): the number of work items lives on the device.
template <int NR, bool AFFINE, bool TRANS>
__global__ void __launch_bounds__(128) lev_diag16_kernel(const BandArgs args, uint32_t *__restrict__ defer,
                                                         const uint32_t *__restrict__ n_dev) {
    const size_t n = n_dev ? (size_t)*n_dev : args.n;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (size_t)gridDim.x * blockDim.x) {
        const size_t pair = args.idx ? (size_t)args.idx[w] : args.pair_base + w;
        const uint64_t a0 = args.a_off[pair], a1 = args.a_off[pair + 1];
        const uint64_t b0 = args.b_off[pair], b1 = args.b_off[pair + 1];
        const uint32_t d = diag16::pair<NR, AFFINE, TRANS>(args.a + a0, a1 - a0, args.b + b0, b1 - b0, args.k, args.mism,
                                                           args.gap, args.sgap, args.tcost);
        if (d == diag16::TOO_WIDE && defer)
            defer[1 + atomicAdd(defer, 1u)] = (uint32_t)pair;
        else
            args.out[pair] = d;
    }
}

template <int NR>
int launch_nr(ta_ctx *ctx, const BandArgs &args, bool affine, bool trans, uint32_t *defer, const uint32_t *n_dev,
              cudaStream_t st) {
    // a thread per work item; with a device-side count the grid is a few waves of persistent threads
    const size_t want = (args.n + 127) / 128;
    const unsigned blocks = (unsigned)(n_dev ? std::min<size_t>(want, (size_t)ctx->sm_count * 4) : want);
    if (affine && trans)
        lev_diag16_kernel<NR, true, true><<<blocks, 128, 0, st>>>(args, defer, n_dev);
    else if (affine)
        lev_diag16_kernel<NR, true, false><<<blocks, 128, 0, st>>>(args, defer, n_dev);
    else if (trans)
        lev_diag16_kernel<NR, false, true><<<blocks, 128, 0, st>>>(args, defer, n_dev);
    else
        lev_diag16_kernel<NR, false, false><<<blocks, 128, 0, st>>>(args, defer, n_dev);
    ctx->launches++;
    TA_CUDA(ctx, cudaGetLastError());
    return TA_OK;
}

int launch_by_width(ta_ctx *ctx, const BandArgs &args, bool affine, bool trans, uint32_t W, uint32_t *defer,
                    const uint32_t *n_dev, cudaStream_t st) {
    if (W <= 8) return launch_nr<2>(ctx, args, affine, trans, defer, n_dev, st);
    if (W <= 12) return launch_nr<3>(ctx, args, affine, trans, defer, n_dev, st);
    if (W <= 16) return launch_nr<4>(ctx, args, affine, trans, defer, n_dev, st);
    if (W <= 24) return launch_nr<6>(ctx, args, affine, trans, defer, n_dev, st);
    return launch_nr<8>(ctx, args, affine, trans, defer, n_dev, st);
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

// The register count follows the band.  W bounds the band of ANY pair of the batch (strings of very different lengths
// need the most diagonals: W = diff + 2 e + 1 with e shrinking as diff grows), W0 is the band of a pair of EQUAL lengths
// (+ 1 for small differences).  When W0 needs fewer registers than W, the batch runs with the smaller count and the
// few pairs whose own band is wider are collected on the device and re-run with the larger one in a second, tiny
// launch (no host round trip: it reads the count from the device) -- 25 % less work per pair on `affine_k16_len128`.
int ta_launch_lev_diag16(ta_ctx *ctx, const BandArgs &args, ta_costs costs, uint32_t W, cudaStream_t st) {
    const bool affine = costs.start_gap != 0, trans = costs.transpose != 0;
    static const bool two_stage = !(getenv("TA_DIAG16_STAGES") && atoi(getenv("TA_DIAG16_STAGES")) == 1);
    const uint32_t kk = args.k;  // (the batch bound already clamped W; W0 only needs k and the costs)
    const uint64_t spare = kk >= 2u * costs.start_gap ? kk - 2u * costs.start_gap : 0u;
    uint64_t W0 = 2 * (spare / (2u * costs.gap)) + 1 + (trans ? 2 : 0) + 1;
    auto regs = [](uint64_t w) { return w <= 8 ? 2 : w <= 12 ? 3 : w <= 16 ? 4 : w <= 24 ? 6 : 8; };
    if (!two_stage || W0 >= W || regs(W0) == regs(W) || args.n > 0xFFFFFFF0ull)
        return launch_by_width(ctx, args, affine, trans, W, nullptr, nullptr, st);
    int rc = ta_dev_reserve(ctx, ctx->d_work[2], (args.n + 1) * sizeof(uint32_t));
    if (rc != TA_OK) return rc;
    uint32_t *defer = (uint32_t *)ctx->d_work[2].p;
    TA_CUDA(ctx, cudaMemsetAsync(defer, 0, sizeof(uint32_t), st));
    if ((rc = launch_by_width(ctx, args, affine, trans, (uint32_t)W0, defer, nullptr, st)) != TA_OK) return rc;
    BandArgs second = args;
    second.idx = defer + 1;
    second.pair_base = 0;
    return launch_by_width(ctx, second, affine, trans, W, nullptr, defer, st);
}
