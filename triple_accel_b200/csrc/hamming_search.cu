// hamming_search.cu -- needle-in-haystack Hamming search for sm_100a (SURVEY.md 8f "next" #2).
//
// Contract: hamming_search_simd_with_opts (reference src/hamming.rs:454-475) == the scalar
// hamming_search_naive_with_opts (src/hamming.rs:96-146) after the public entry's checks: needle longer than the
// haystack or empty needle -> no matches; a NUL byte in the haystack -> panic (src/lib.rs:237-243, reported as
// TA_ERR_NUL_BYTE).  Every start position i with mismatches(needle, haystack[i .. i+N)) <= k is a hit
// Match{start: i, end: i + N, k: mismatches}; SearchType::Best shrinks the threshold to the running minimum while
// scanning (:124-127) and keeps the hits equal to the final minimum (:136-143) -- applied on the host to the
// sparse hit list.
//
// Kernel: one thread per start position (grid-stride over all positions of all haystacks via a per-haystack
// prefix of position counts), needle in shared memory (broadcast reads), haystack bytes through the read-only
// path: consecutive threads read consecutive bytes, so each 128-byte line is fetched from HBM once and served
// N times from L1.  Algorithmic bytes: |haystack|.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "ta_common.cuh"

namespace {

struct HHit {
    uint32_t hay, cost;
    uint64_t start;
};

template <bool SMEM>  // needle staged in shared memory (the rule), or read through L1 / L2 when it does not fit
__global__ void __launch_bounds__(256) hamming_search_kernel(const uint8_t *__restrict__ needle, uint32_t N,
                                                             const uint8_t *__restrict__ hay,
                                                             const uint64_t *__restrict__ hay_off, size_t n,
                                                             uint64_t max_hay, uint32_t k, HHit *__restrict__ hits,
                                                             unsigned long long *__restrict__ hit_count,
                                                             unsigned long long hit_cap, uint32_t *__restrict__ nul_flag) {
    extern __shared__ uint8_t sneedle_buf[];
    const uint8_t *sneedle = needle;
    if (SMEM) {
        for (uint32_t q = threadIdx.x; q < N; q += blockDim.x) sneedle_buf[q] = needle[q];
        __syncthreads();
        sneedle = sneedle_buf;
    }
    // blockIdx.y walks haystacks (grid-stride), blockIdx.x * blockDim.x + threadIdx.x walks positions
    for (size_t h = blockIdx.y; h < n; h += gridDim.y) {
        const uint64_t h0 = hay_off[h], H = hay_off[h + 1] - h0;
        if (N > H) continue;  // src/hamming.rs:455-457
        const uint8_t *p = hay + h0;
        const uint64_t len = H + 1 - N;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < H; i += (uint64_t)gridDim.x * blockDim.x) {
            const uint32_t c0 = __ldg(p + i);
            if (c0 == 0) atomicExch(nul_flag, 1u);  // check_no_null_bytes covers the whole haystack
            if (i >= len) continue;
            uint32_t cnt = c0 != sneedle[0];
            for (uint32_t j = 1; j < N && cnt <= k; j++) cnt += __ldg(p + i + j) != sneedle[j];
            if (cnt <= k) {
                const unsigned long long slot = atomicAdd(hit_count, 1ull);
                if (slot < hit_cap) hits[slot] = HHit{(uint32_t)h, cnt, i};
            }
        }
    }
    (void)max_hay;
}

}  // namespace

static int hamming_search_impl(ta_ctx *ctx, bool check_nul, const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                               const uint64_t *hay_off, size_t n, uint32_t k, int search_type, ta_match **out_matches,
                               uint64_t **out_match_off);

extern "C" int ta_hamming_search_batch(ta_ctx *ctx, const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                                       const uint64_t *hay_off, size_t n, uint32_t k, int search_type,
                                       ta_match **out_matches, uint64_t **out_match_off) {
    return hamming_search_impl(ctx, true, needle, needle_len, hay, hay_off, n, k, search_type, out_matches, out_match_off);
}

// hamming_search_naive_with_opts (src/hamming.rs:96-146): the scalar routine has no NUL-byte restriction (the check
// lives in the SIMD entry, src/hamming.rs:463); same kernel, the NUL flag is simply not an error.
extern "C" int ta_hamming_search_naive_batch(ta_ctx *ctx, const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                                             const uint64_t *hay_off, size_t n, uint32_t k, int search_type,
                                             ta_match **out_matches, uint64_t **out_match_off) {
    return hamming_search_impl(ctx, false, needle, needle_len, hay, hay_off, n, k, search_type, out_matches, out_match_off);
}

static int hamming_search_impl(ta_ctx *ctx, bool check_nul, const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                               const uint64_t *hay_off, size_t n, uint32_t k, int search_type, ta_match **out_matches,
                               uint64_t **out_match_off) {
    if (!ctx || !out_matches || !out_match_off) return TA_ERR_BAD_ARG;
    *out_matches = nullptr;
    *out_match_off = nullptr;
    if (search_type != TA_SEARCH_ALL && search_type != TA_SEARCH_BEST) return TA_ERR_BAD_ARG;
    if (n && !hay_off) return TA_ERR_BAD_ARG;
    if (needle_len && !needle) return TA_ERR_BAD_ARG;
    if (n > 0xFFFFFFF0ull || needle_len > TA_MAX_STRING_LEN) return TA_ERR_TOO_LARGE;
    uint64_t max_hay = 0, total_hay = 0;
    for (size_t i = 0; i < n; i++) {
        if (hay_off[i + 1] < hay_off[i]) return TA_ERR_BAD_ARG;
        max_hay = std::max(max_hay, hay_off[i + 1] - hay_off[i]);
    }
    if (n) total_hay = hay_off[n] - hay_off[0];
    if (total_hay && !hay) return TA_ERR_BAD_ARG;
    if (ctx->multi && n) {  // several GPUs: ranges of haystacks balanced by bytes, match lists concatenated in range order
        std::lock_guard<std::mutex> lock(ctx->mu);
        const int parts = ta_multi_parts(ctx, total_hay, n);
        std::vector<size_t> bound;
        ta_multi_bounds(hay_off, nullptr, n, parts, bound);
        std::vector<ta_match *> ms(parts, nullptr);
        std::vector<uint64_t *> mo(parts, nullptr);
        const int rc = ta_multi_run(ctx, parts, [&](int r) -> int {
            const size_t lo = bound[r], cnt = bound[r + 1] - bound[r];
            if (cnt == 0) return TA_OK;
            return hamming_search_impl(ta_multi_sub(ctx, r), check_nul, needle, needle_len, hay, hay_off + lo, cnt, k,
                                       search_type, &ms[r], &mo[r]);
        });
        if (rc != TA_OK) {
            for (int r = 0; r < parts; r++) ta_free(ms[r]), ta_free(mo[r]);
            return rc;
        }
        return ta_concat_lists<ta_match>(parts, bound, n, ms, mo, out_matches, out_match_off);
    }
    uint64_t *moff = (uint64_t *)ta_out_alloc((n + 1) * sizeof(uint64_t));
    if (moff) memset(moff, 0, (n + 1) * sizeof(uint64_t));
    if (!moff) return TA_ERR_NOMEM;
    std::vector<ta_match> result;
    auto finish = [&]() {
        ta_match *m = (ta_match *)ta_out_alloc((result.size() ? result.size() : 1) * sizeof(ta_match));
        if (!m) {
            ta_free(moff);
            return (int)TA_ERR_NOMEM;
        }
        if (!result.empty()) memcpy(m, result.data(), result.size() * sizeof(ta_match));
        *out_matches = m;
        *out_match_off = moff;
        return (int)TA_OK;
    };
    if (needle_len == 0 || n == 0 || needle_len > max_hay) return finish();  // src/hamming.rs:455-461

    std::vector<HHit> hits;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        auto run = [&]() -> int {
            TA_CUDA(ctx, cudaSetDevice(ctx->device));
            cudaStream_t st = ctx->stream;
            int rc;
            const uint64_t lo = hay_off[0];
            if ((rc = ta_dev_reserve(ctx, ctx->d_a[0], total_hay + 64)) != TA_OK) return rc;
            if ((rc = ta_dev_reserve(ctx, ctx->d_aoff[0], (n + 1) * sizeof(uint64_t))) != TA_OK) return rc;
            if ((rc = ta_dev_reserve(ctx, ctx->d_b[0], needle_len + 64)) != TA_OK) return rc;
            if (total_hay)
                TA_CUDA(ctx, cudaMemcpyAsync(ctx->d_a[0].p, hay + lo, total_hay, cudaMemcpyHostToDevice, st));
            TA_CUDA(ctx, cudaMemcpyAsync(ctx->d_aoff[0].p, hay_off, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
            TA_CUDA(ctx, cudaMemcpyAsync(ctx->d_b[0].p, needle, needle_len, cudaMemcpyHostToDevice, st));
            const uint8_t *d_hay = (const uint8_t *)ctx->d_a[0].p - lo;
            uint32_t *nul_flag = ctx->d_flags + 8;
            unsigned long long *d_count = (unsigned long long *)(ctx->d_flags + 4);
            unsigned long long cap = std::max<unsigned long long>(4096, n * 8);
            const unsigned gx = (unsigned)std::min<uint64_t>((max_hay + 255) / 256, 4096);
            const unsigned gy = (unsigned)std::min<size_t>(n, 65535);
            for (int attempt = 0; attempt < 2; attempt++) {
                if ((rc = ta_dev_reserve(ctx, ctx->d_work[1], cap * sizeof(HHit))) != TA_OK) return rc;
                TA_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
                TA_CUDA(ctx, cudaMemsetAsync(nul_flag, 0, sizeof(uint32_t), st));
                // needles above 48 KB opt in to more dynamic shared memory; above the device's limit (227 KB) the kernel
                // reads the needle from global memory (such needles used to fail at launch)
                const size_t nsmem = (needle_len + 15) & ~(size_t)15;
                if (nsmem <= (size_t)ctx->smem_optin) {
                    if (nsmem > 48 * 1024)
                        TA_CUDA(ctx, cudaFuncSetAttribute(hamming_search_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nsmem));
                    hamming_search_kernel<true><<<dim3(gx, gy), 256, nsmem, st>>>(
                        (const uint8_t *)ctx->d_b[0].p, (uint32_t)needle_len, d_hay, (const uint64_t *)ctx->d_aoff[0].p, n,
                        max_hay, k, (HHit *)ctx->d_work[1].p, d_count, cap, nul_flag);
                } else {
                    hamming_search_kernel<false><<<dim3(gx, gy), 256, 0, st>>>(
                        (const uint8_t *)ctx->d_b[0].p, (uint32_t)needle_len, d_hay, (const uint64_t *)ctx->d_aoff[0].p, n,
                        max_hay, k, (HHit *)ctx->d_work[1].p, d_count, cap, nul_flag);
                }
                ctx->launches++;
                TA_CUDA(ctx, cudaGetLastError());
                TA_CUDA(ctx, cudaMemcpyAsync(ctx->h_flags + 4, ctx->d_flags + 4, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                TA_CUDA(ctx, cudaStreamSynchronize(st));
                if (check_nul && ctx->h_flags[8]) return TA_ERR_NUL_BYTE;
                unsigned long long got;
                memcpy(&got, ctx->h_flags + 4, sizeof got);
                if (got <= cap) {
                    hits.resize((size_t)got);
                    if (got)
                        TA_CUDA(ctx, cudaMemcpyAsync(hits.data(), ctx->d_work[1].p, (size_t)got * sizeof(HHit),
                                                     cudaMemcpyDeviceToHost, st));
                    TA_CUDA(ctx, cudaStreamSynchronize(st));
                    return TA_OK;
                }
                cap = got;
            }
            return TA_ERR_TOO_LARGE;
        };
        const int rc = run();
        if (rc != TA_OK) {
            cudaStreamSynchronize(ctx->stream);
            ta_free(moff);
            return rc;
        }
    }
    std::sort(hits.begin(), hits.end(), [](const HHit &x, const HHit &y) {
        return x.hay != y.hay ? x.hay < y.hay : x.start < y.start;
    });
    const bool best = search_type == TA_SEARCH_BEST;
    size_t hp = 0;
    std::vector<ta_match> cur;
    for (size_t i = 0; i < n; i++) {
        cur.clear();
        uint32_t curr_k = k;
        for (; hp < hits.size() && hits[hp].hay == i; hp++) {
            const HHit &h = hits[hp];
            if (h.cost <= curr_k) {  // src/hamming.rs:112-131 (the early stop only skips positions above curr_k)
                if (best) curr_k = h.cost;
                cur.push_back(ta_match{h.start, h.start + needle_len, h.cost, 0});
            }
        }
        if (best) {  // src/hamming.rs:136-143
            size_t f = 0;
            for (size_t r = 0; r < cur.size(); r++)
                if (cur[r].k == curr_k) cur[f++] = cur[r];
            cur.resize(f);
        }
        result.insert(result.end(), cur.begin(), cur.end());
        moff[i + 1] = result.size();
    }
    return finish();
}
