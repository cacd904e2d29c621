// ta_common.cuh -- shared declarations for the sm_100a kernels and the C-ABI host layer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <functional>
#include <mutex>
#include <string>
#include <vector>

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
    DevBuf d_a[2], d_b[2], d_aoff[2], d_boff[2], d_out[2], d_work[5];
    DevBuf h_pin[4];         // pinned staging for pageable inputs / outputs
    int len_hint = -1;          // ta_set_length_hint: -1 auto, 0 equal lengths, 1 ragged
    bool batch_ragged = false;  // decided per batch by the entry points, read by the unit-cost dispatcher
    uint32_t *d_flags = nullptr;  // [0] = deferred error code of *_dev kernels, [1..] scratch counters
    uint32_t *h_flags = nullptr;  // pinned mirror
    uint64_t launches = 0;
    std::string last_error;
    struct ta_multi *multi = nullptr;  // non-null: this ctx spans several devices (multi.cu) and owns no CUDA state itself
};

// grow-only device / pinned-host buffers
int ta_dev_reserve(ta_ctx *ctx, DevBuf &b, size_t bytes);
int ta_pin_reserve(ta_ctx *ctx, DevBuf &b, size_t bytes);
int ta_cuda_fail(ta_ctx *ctx, cudaError_t e, const char *what);
// caller-owned output arrays (released with ta_free); freed blocks are cached for reuse, see api.cu
extern "C" void *ta_out_alloc(size_t bytes);

// ---- multi-device contexts (multi.cu) ----------------------------------------------------------------------------
int ta_multi_size(ta_ctx *ctx);                                    // number of devices (1 for a plain ctx)
ta_ctx *ta_multi_sub(ta_ctx *ctx, int r);                          // single-device sub-context r
int ta_multi_parts(ta_ctx *ctx, uint64_t total_bytes, size_t n);   // how many shards a batch of this size is cut into
// bound[r] .. bound[r + 1] = units of shard r, balanced by bytes (b_off may be null)
void ta_multi_bounds(const uint64_t *a_off, const uint64_t *b_off, size_t n, int parts, std::vector<size_t> &bound);
// fn(r) for every shard, shard 0 on the calling thread and the others on their device's worker thread; first error wins
int ta_multi_run(ta_ctx *ctx, int parts, const std::function<int(int)> &fn);
// needle -> ctx->d_b[0] of the first `parts` sub-contexts (ncclBroadcast from device 0 when all devices take part)
int ta_multi_needle(ta_ctx *ctx, const uint8_t *needle, size_t needle_len, int parts);
void ta_multi_shutdown(ta_ctx *ctx);

// Concatenates per-shard (items, offsets) outputs -- shard r covers units bound[r] .. bound[r + 1] and its arrays came
// from ta_out_alloc -- into one pair of caller-owned arrays; frees the shard arrays.
// q-gram search filter: queue capacity = haystack bytes / this + 4096 entries (lev_bitpar.cu); search.cu sizes the
// exact kernel's item list from it (<= 8 distinct 16-byte granules per queue entry)
#define TA_QGRAM_BYTES_PER_ENTRY 256

// what ta_launch_search_filter (lev_bitpar.cu) queued: a device list of codes haystack * segs + granule, `gran` bytes of
// end positions per granule.  If alt_flag is non-null and set ON THE DEVICE when the exact kernel runs, the list was
// written by the fallback filter instead and holds alt_segs / alt_gran codes.
struct ta_filter_out {
    uint32_t segs = 0, gran = 0;
    const uint32_t *alt_flag = nullptr;
    uint32_t alt_segs = 0, alt_gran = 0;
};

template <typename T>
int ta_concat_lists(int parts, const std::vector<size_t> &bound, size_t n, std::vector<T *> &items,
                    std::vector<uint64_t *> &offs, T **out_items, uint64_t **out_off) {
    uint64_t total = 0;
    for (int r = 0; r < parts; r++)
        if (offs[r]) total += offs[r][bound[r + 1] - bound[r]];
    uint64_t *off = (uint64_t *)ta_out_alloc((n + 1) * sizeof(uint64_t));
    T *it = (T *)ta_out_alloc((total ? total : 1) * sizeof(T));
    int rc = (off && it) ? TA_OK : TA_ERR_NOMEM;
    if (rc == TA_OK) {
        uint64_t base = 0;
        off[0] = 0;
        for (int r = 0; r < parts; r++) {
            const size_t cnt = bound[r + 1] - bound[r];
            if (!offs[r]) {  // an empty shard
                for (size_t i = 0; i < cnt; i++) off[bound[r] + i + 1] = base;
                continue;
            }
            for (size_t i = 0; i < cnt; i++) off[bound[r] + i + 1] = base + offs[r][i + 1];
            if (offs[r][cnt]) memcpy(it + base, items[r], offs[r][cnt] * sizeof(T));
            base += offs[r][cnt];
        }
        *out_items = it;
        *out_off = off;
    } else {
        ta_free(off);
        ta_free(it);
    }
    for (int r = 0; r < parts; r++) {
        ta_free(items[r]);
        ta_free(offs[r]);
    }
    return rc;
}


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
// Diagonal-extension kernel (unit costs, long strings, min(k, max_len) <= 64), lev_fr.cu
bool ta_fr_can_handle(uint32_t k, ta_costs c, uint32_t max_len);
bool ta_fr_preferred(uint32_t k, ta_costs c, uint32_t max_len);
int ta_launch_lev_fr(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                     size_t n, const uint32_t *idx, uint32_t k, ta_costs costs, uint32_t max_len, uint32_t *out,
                     cudaStream_t st);
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
