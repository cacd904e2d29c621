// hamming.cu -- batched Hamming distance for sm_100a.
//
// Contract: hamming(a, b) == hamming_naive(a, b) (reference src/hamming.rs:36-47, 390-392): the number of
// positions where the two equal-length strings differ.  The reference's AVX2 path (src/jewel.rs:2320-2365) walks
// 32 bytes per step; here a group of G lanes (G picked from the mean string length) walks one pair in 16-byte
// vector loads, compares with the byte-SIMD __vcmpne4 + popcount, and reduces with shuffles.  The kernel is
// HBM-bound: algorithmic traffic is |a| + |b| + 4 bytes per pair (+16 B of CSR offsets).
#include "ta_common.cuh"

namespace {

__device__ __forceinline__ uint32_t ne_bytes(uint32_t x, uint32_t y) {
    // __vcmpne4 gives 0xff per differing byte; popc/8 = number of differing bytes
    return __popc(__vcmpne4(x, y)) >> 3;
}

template <int G>
__global__ void __launch_bounds__(256) hamming_kernel(const uint8_t *__restrict__ a, const uint64_t *__restrict__ a_off,
                                                      const uint8_t *__restrict__ b, const uint64_t *__restrict__ b_off,
                                                      size_t n, uint32_t *__restrict__ out,
                                                      uint32_t *__restrict__ err_flag) {
    const unsigned lane = threadIdx.x & (G - 1);
    const size_t groups_per_grid = (size_t)gridDim.x * (blockDim.x / G);
    const unsigned full = 0xffffffffu;

    for (size_t pair = (size_t)blockIdx.x * (blockDim.x / G) + threadIdx.x / G;; pair += groups_per_grid) {
        // keep whole warps in the loop so the shuffles below are always convergent
        const size_t warp_first = pair - (threadIdx.x & 31) / G;
        if (warp_first >= n) break;
        const bool live = pair < n;
        uint32_t cnt = 0;
        bool bad = false;
        if (live) {
            const uint64_t a0 = a_off[pair], a1 = a_off[pair + 1];
            const uint64_t b0 = b_off[pair], b1 = b_off[pair + 1];
            const uint64_t len = a1 - a0;
            bad = (b1 - b0) != len;
            if (!bad) {
                const uint8_t *pa = a + a0, *pb = b + b0;
                const uintptr_t ua = (uintptr_t)pa, ub = (uintptr_t)pb;
                if (((ua | ub) & 15) == 0) {
                    const uint4 *va = (const uint4 *)pa, *vb = (const uint4 *)pb;
                    const uint64_t nv = len >> 4;
                    for (uint64_t q = lane; q < nv; q += G) {
                        const uint4 x = __ldg(va + q), y = __ldg(vb + q);
                        cnt += ne_bytes(x.x, y.x) + ne_bytes(x.y, y.y) + ne_bytes(x.z, y.z) + ne_bytes(x.w, y.w);
                    }
                    for (uint64_t q = (nv << 4) + lane; q < len; q += G) cnt += pa[q] != pb[q];
                } else if (((ua ^ ub) & 3) == 0) {
                    // same misalignment modulo 4: byte head, u32 body, byte tail
                    uint64_t head = (4 - (ua & 3)) & 3;
                    if (head > len) head = len;
                    for (uint64_t q = lane; q < head; q += G) cnt += pa[q] != pb[q];
                    const uint32_t *wa = (const uint32_t *)(pa + head), *wb = (const uint32_t *)(pb + head);
                    const uint64_t nw = (len - head) >> 2;
                    for (uint64_t q = lane; q < nw; q += G) cnt += ne_bytes(__ldg(wa + q), __ldg(wb + q));
                    for (uint64_t q = head + (nw << 2) + lane; q < len; q += G) cnt += pa[q] != pb[q];
                } else {
                    for (uint64_t q = lane; q < len; q += G) cnt += pa[q] != pb[q];
                }
            }
        }
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) cnt += __shfl_xor_sync(full, cnt, o, G);
        if (live && lane == 0) {
            if (bad) {
                out[pair] = TA_NONE;
                atomicExch(err_flag, (uint32_t)(-TA_ERR_LEN_MISMATCH));
            } else {
                out[pair] = cnt;
            }
        }
    }
}

template <int G>
void launch(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off, size_t n,
            uint32_t *out, uint32_t *err_flag, cudaStream_t st) {
    const int threads = 256;
    const size_t groups_per_block = threads / G;
    size_t blocks = (n + groups_per_block - 1) / groups_per_block;
    const size_t cap = (size_t)ctx->sm_count * 8 * 4;  // 8 resident CTAs/SM, up to 4 waves; grid-stride beyond that
    if (blocks > cap) blocks = cap;
    hamming_kernel<G><<<(unsigned)blocks, threads, 0, st>>>(a, a_off, b, b_off, n, out, err_flag);
}

}  // namespace

int ta_launch_hamming(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                      size_t n, uint32_t avg_len, uint32_t *out, uint32_t *err_flag, cudaStream_t st) {
    if (n == 0) return TA_OK;
    const uint32_t chunks = (avg_len + 15) / 16;  // 16-byte vectors per string
    if (chunks <= 1)
        launch<1>(ctx, a, a_off, b, b_off, n, out, err_flag, st);
    else if (chunks <= 2)
        launch<2>(ctx, a, a_off, b, b_off, n, out, err_flag, st);
    else if (chunks <= 4)
        launch<4>(ctx, a, a_off, b, b_off, n, out, err_flag, st);
    else if (chunks <= 8)
        launch<8>(ctx, a, a_off, b, b_off, n, out, err_flag, st);
    else if (chunks <= 16)
        launch<16>(ctx, a, a_off, b, b_off, n, out, err_flag, st);
    else
        launch<32>(ctx, a, a_off, b, b_off, n, out, err_flag, st);
    ctx->launches++;
    TA_CUDA(ctx, cudaGetLastError());
    return TA_OK;
}
