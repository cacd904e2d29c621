// lev_bitpar.cu -- bit-parallel fast paths for unit costs (LEVENSHTEIN_COSTS / RDAMERAU_COSTS) on sm_100a.
//
//  * lev_bitpar32_kernel: k-bounded distance, one thread per pair, band of <= 32 diagonals (k <= 31, or <= 29 with
//    transpositions).  Algorithm and data movement are described in lev_bitpar_core.cuh.  Same contract as the
//    general kernel in lev_band.cu (reference src/levenshtein.rs:376-545); the dispatcher in api.cu picks this one
//    whenever the cost model is unit and the band fits.
//  * search_filter_kernel: (below) flags haystacks that contain at least one match end with cost <= k.
#include "lev_bitpar_core.cuh"
#include "ta_common.cuh"

namespace {

template <bool TRANS>
__global__ void __launch_bounds__(128) lev_bitpar32_kernel(const uint8_t *__restrict__ a,
                                                           const uint64_t *__restrict__ a_off,
                                                           const uint8_t *__restrict__ b,
                                                           const uint64_t *__restrict__ b_off,
                                                           const uint32_t *__restrict__ idx, size_t n, uint32_t k,
                                                           uint32_t *__restrict__ out) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const size_t pair = idx ? (size_t)idx[w] : w;
    const uint64_t a0 = a_off[pair], a1 = a_off[pair + 1];
    const uint64_t b0 = b_off[pair], b1 = b_off[pair + 1];
    out[pair] = bitpar::pair_unit_costs<TRANS>(a + a0, a1 - a0, b + b0, b1 - b0, k);
}

}  // namespace

bool ta_bitpar_can_handle(uint32_t k, ta_costs c, uint32_t max_len) {
    if (!(c.mismatch == 1 && c.gap == 1 && c.start_gap == 0 && c.transpose <= 1)) return false;
    const uint32_t kk = k < max_len ? k : max_len;  // max_k = min(k, n) for unit costs
    return kk <= (c.transpose ? 29u : 31u);
}

int ta_launch_lev_bitpar(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                         const uint64_t *b_off, size_t n, const uint32_t *idx, uint32_t k, ta_costs costs,
                         uint32_t *out, cudaStream_t st) {
    if (n == 0) return TA_OK;
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (costs.transpose)
        lev_bitpar32_kernel<true><<<blocks, 128, 0, st>>>(a, a_off, b, b_off, idx, n, k, out);
    else
        lev_bitpar32_kernel<false><<<blocks, 128, 0, st>>>(a, a_off, b, b_off, idx, n, k, out);
    ctx->launches++;
    TA_CUDA(ctx, cudaGetLastError());
    return TA_OK;
}

int ta_launch_search_filter(ta_ctx *ctx, const uint8_t *needle_dev, uint32_t needle_len, const uint8_t *hay,
                            const uint64_t *hay_off, size_t n, uint32_t k, bool transpose, uint32_t *idx_out,
                            uint32_t *counter, cudaStream_t st) {
    (void)ctx, (void)needle_dev, (void)needle_len, (void)hay, (void)hay_off, (void)n, (void)k, (void)transpose;
    (void)idx_out, (void)counter, (void)st;
    return TA_ERR_TOO_LARGE;  // not available: caller runs the exact kernel on every haystack
}
