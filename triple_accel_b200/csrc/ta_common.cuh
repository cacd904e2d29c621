// ta_common.cuh -- shared declarations for the sm_100a kernels and the C-ABI host layer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

#include "../../include/triple_accel_b200.h"

#define TA_SEARCH_SEG 512  // haystack bytes scanned per pre-filter thread (lev_bitpar.cu)
#define TA_SEARCH_SUB 128  // granularity at which the pre-filter flags match ends = work item of the exact kernel
#define TA_INF 0x3FFFFFFFu  // "out of band" cell value; real costs stay below 2^30 (TA_MAX_STRING_LEN)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct ta_ctx {
    int device = 0;
    int sm_count = 148;
    int smem_optin = 0;  // max opt-in dynamic shared memory per block
    std::mutex mu;
    cudaStream_t stream = nullptr;   // compute + D2H
    cudaStream_t stream2 = nullptr;  // H2D of the next chunk
    static constexpr int MAX_CHUNKS = 8;
    cudaEvent_t ev_h2d[MAX_CHUNKS] = {};  // "chunk c is in device memory"
    DevBuf d_a[2], d_b[2], d_aoff[2], d_boff[2], d_out[2], d_work[4];
    DevBuf h_pin[4];         // pinned staging for pageable inputs / outputs
    uint32_t *d_flags = nullptr;  // [0] = deferred error code of *_dev kernels, [1..] scratch counters
    uint32_t *h_flags = nullptr;  // pinned mirror
    uint64_t launches = 0;
    std::string last_error;
};

// grow-only device / pinned-host buffers
int ta_dev_reserve(ta_ctx *ctx, DevBuf &b, size_t bytes);
int ta_pin_reserve(ta_ctx *ctx, DevBuf &b, size_t bytes);
int ta_cuda_fail(ta_ctx *ctx, cudaError_t e, const char *what);
// caller-owned output arrays (released with ta_free); freed blocks are cached for reuse, see api.cu
extern "C" void *ta_out_alloc(size_t bytes);

#define TA_CUDA(ctx, call)                                        \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return ta_cuda_fail(ctx, e__, #call); \
    } while (0)

// ---- kernel launchers (device pointers; asynchronous on `st`) --------------------------------------------------
int ta_launch_hamming(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                      size_t n, uint32_t avg_len, uint32_t *out, uint32_t *err_flag, cudaStream_t st);

// General banded anti-diagonal DP (all cost models, any k).  `idx` (may be null) is an indirection: work item w
// processes pair idx[w] (used by the exponential-k driver to re-run only the pairs that are still TA_NONE).
int ta_launch_lev_band(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                       size_t n, const uint32_t *idx, uint32_t k, ta_costs costs, uint32_t max_len, uint32_t *out,
                       cudaStream_t st);

// Bit-parallel fast path (unit costs, band <= 32 diagonals), lev_bitpar.cu
bool ta_bitpar_can_handle(uint32_t k, ta_costs c, uint32_t max_len);
int ta_launch_lev_bitpar(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                         const uint64_t *b_off, size_t n, const uint32_t *idx, uint32_t k, ta_costs costs,
                         uint32_t max_len, uint32_t *out, cudaStream_t st);
// dispatcher: bit-parallel kernel when the cost model is unit and the band fits, else the general banded kernel
int ta_launch_lev(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                  size_t n, const uint32_t *idx, uint32_t k, ta_costs costs, uint32_t max_len, uint32_t *out,
                  cudaStream_t st);

// Traceback (trace_on = true): TRACE variant of the banded kernel + walk kernels, lev_band.cu
uint32_t ta_trace_cells(uint32_t W);
int ta_launch_lev_band_trace(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                             const uint64_t *b_off, size_t n, const uint32_t *idx, size_t pair_base, uint32_t k,
                             ta_costs costs, uint32_t max_len, uint32_t *out, uint8_t *trace, size_t trace_stride,
                             cudaStream_t st);
int ta_launch_trace_walk(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                         size_t n, const uint32_t *idx, size_t pair_base, uint32_t k, ta_costs costs, uint32_t max_len,
                         const uint32_t *out, const uint8_t *trace, size_t trace_stride, uint32_t *counts,
                         const uint64_t *edit_off, ta_edit *edits, cudaStream_t st);

// band width (number of diagonals) the general kernel needs in the worst case for (k, costs, max_len)
uint32_t ta_band_width_bound(uint32_t k, ta_costs c, uint32_t max_len);
