// api.cu -- C-ABI host layer of libtriple_accel_b200.so (see include/triple_accel_b200.h).
// Device memory, streams and staging live in ta_ctx; every compute entry point runs CUDA kernels and fails with
// TA_ERR_CUDA when no device is usable -- there is no CPU path in this library.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "ta_common.cuh"

int ta_cuda_fail(ta_ctx *ctx, cudaError_t e, const char *what) {
    if (ctx) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
        ctx->last_error = buf;
    }
    (void)cudaGetLastError();  // clear the sticky-less error state
    return TA_ERR_CUDA;
}

int ta_launch_lev(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                  size_t n, const uint32_t *idx, uint32_t k, ta_costs costs, uint32_t max_len, uint32_t *out,
                  cudaStream_t st) {
    static const bool force_band = getenv("TA_FORCE_BAND") != nullptr;  // testing: exercise the general kernel
    if (!force_band && ta_fr_preferred(k, costs, max_len))  // long strings, small k: diagonal extension (lev_fr.cu)
        return ta_launch_lev_fr(ctx, a, a_off, b, b_off, n, idx, k, costs, max_len, out, st);
    if (!force_band && ta_bitpar_can_handle(k, costs, max_len))
        return ta_launch_lev_bitpar(ctx, a, a_off, b, b_off, n, idx, k, costs, max_len, out, st);
    return ta_launch_lev_band(ctx, a, a_off, b, b_off, n, idx, k, costs, max_len, out, st);
}

int ta_dev_reserve(ta_ctx *ctx, DevBuf &b, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes <= b.cap && b.p) return TA_OK;
    if (b.p) TA_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 4 + 256;  // headroom: batches of similar size do not reallocate
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        want = bytes + 256;
        TA_CUDA(ctx, cudaMalloc(&b.p, want));
    }
    b.cap = want;
    return TA_OK;
}

int ta_pin_reserve(ta_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap && b.p) return TA_OK;
    if (b.p) TA_CUDA(ctx, cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
    TA_CUDA(ctx, cudaMallocHost(&b.p, bytes + 256));
    b.cap = bytes + 256;
    return TA_OK;
}

static void destroy_ctx_resources(ta_ctx *ctx);

extern "C" {

int ta_abi_version(void) { return TA_ABI_VERSION; }

const char *ta_strerror(int code) {
    switch (code) {
        case TA_OK: return "ok";
        case TA_ERR_CUDA: return "CUDA error (see ta_last_error)";
        case TA_ERR_LEN_MISMATCH: return "hamming: strings of a pair differ in length";
        case TA_ERR_BAD_COSTS: return "invalid EditCosts";
        case TA_ERR_BAD_ARG: return "bad argument";
        case TA_ERR_TOO_LARGE: return "input too large for this build";
        case TA_ERR_NOMEM: return "out of host memory";
        case TA_ERR_NUL_BYTE: return "hamming_search: zero/null bytes are not allowed in the haystack";
        default: return "unknown error";
    }
}

int ta_init(int device, ta_ctx **out) {
    if (!out) return TA_ERR_BAD_ARG;
    *out = nullptr;
    ta_ctx *ctx = new (std::nothrow) ta_ctx();
    if (!ctx) return TA_ERR_NOMEM;
    ctx->device = device;
    auto fail = [&](cudaError_t e, const char *what) {
        ta_cuda_fail(ctx, e, what);
        fprintf(stderr, "triple_accel_b200: %s\n", ctx->last_error.c_str());
        destroy_ctx_resources(ctx);  // whatever was created before the failure
        delete ctx;
        return TA_ERR_CUDA;
    };
    cudaError_t e;
    int count = 0;
    if ((e = cudaGetDeviceCount(&count)) != cudaSuccess) return fail(e, "cudaGetDeviceCount");
    if (device < 0 || device >= count) return fail(cudaErrorInvalidDevice, "ta_init(device)");
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(e, "cudaSetDevice");
    if ((e = cudaFree(0)) != cudaSuccess) return fail(e, "cudaFree(0): no usable CUDA context");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(e, "cudaGetDeviceProperties");
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    if ((e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    for (int i = 0; i < ta_ctx::MAX_CHUNKS; i++)
        if ((e = cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming)) != cudaSuccess) return fail(e, "event");
    if ((e = cudaMalloc((void **)&ctx->d_flags, 64 * sizeof(uint32_t))) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMemset(ctx->d_flags, 0, 64 * sizeof(uint32_t))) != cudaSuccess) return fail(e, "cudaMemset");
    if ((e = cudaMallocHost((void **)&ctx->h_flags, 64 * sizeof(uint32_t))) != cudaSuccess) return fail(e, "cudaMallocHost");
    memset(ctx->h_flags, 0, 64 * sizeof(uint32_t));
    *out = ctx;
    return TA_OK;
}

void ta_shutdown(ta_ctx *ctx) {
    if (!ctx) return;
    if (ctx->multi) {  // a multi-device ctx owns sub-contexts, worker threads and the NCCL communicator only
        ta_multi_shutdown(ctx);
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    destroy_ctx_resources(ctx);
    delete ctx;
}

}  // extern "C"

static void destroy_ctx_resources(ta_ctx *ctx) {
    DevBuf *dev[] = {&ctx->d_a[0], &ctx->d_a[1], &ctx->d_b[0], &ctx->d_b[1], &ctx->d_aoff[0], &ctx->d_aoff[1],
                     &ctx->d_boff[0], &ctx->d_boff[1], &ctx->d_out[0], &ctx->d_out[1], &ctx->d_work[0],
                     &ctx->d_work[1], &ctx->d_work[2], &ctx->d_work[3], &ctx->d_work[4]};
    for (DevBuf *b : dev)
        if (b->p) cudaFree(b->p);
    for (DevBuf &b : ctx->h_pin)
        if (b.p) cudaFreeHost(b.p);
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    for (int i = 0; i < ta_ctx::MAX_CHUNKS; i++)
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    (void)cudaGetLastError();
}

extern "C" {

const char *ta_last_error(ta_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
int ta_device(ta_ctx *ctx) { return ctx ? ctx->device : -1; }
uint64_t ta_launch_count(ta_ctx *ctx) { return ctx ? ctx->launches : 0; }

void *ta_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return p;
}
void ta_host_free(void *p) {
    if (p) cudaFreeHost(p);
}
// Output arrays (match lists, per-unit offsets, edit lists) are handed to the caller and come back through ta_free.
// Fresh multi-hundred-KB mallocs are mmap'ed and page-faulted in on first touch (~1 us per 4 KB page: 0.2 ms for the
// offsets of 100 k haystacks), so freed blocks are parked in a small cache and reused by the next call of similar size.
namespace {
struct OutHdr {
    uint64_t magic, cap;
};
constexpr uint64_t OUT_MAGIC = 0x74615f6f75745f31ull;
constexpr uint64_t OUT_PARKED = 0x74615f7061726b31ull;  // in the cache: a second ta_free of it is ignored
constexpr int OUT_SLOTS = 6;
constexpr uint64_t OUT_MAX_PARKED = 64ull << 20;  // larger blocks go straight back to the allocator
std::mutex g_out_mu;
OutHdr *g_out_cache[OUT_SLOTS] = {};
}  // namespace

void *ta_out_alloc(size_t bytes) {
    if (bytes == 0) bytes = 1;
    {
        std::lock_guard<std::mutex> lock(g_out_mu);
        for (int i = 0; i < OUT_SLOTS; i++) {
            OutHdr *h = g_out_cache[i];
            if (h && h->cap >= bytes && h->cap <= 2 * bytes + (64u << 10)) {
                g_out_cache[i] = nullptr;
                h->magic = OUT_MAGIC;
                return h + 1;
            }
        }
    }
    OutHdr *h = (OutHdr *)malloc(sizeof(OutHdr) + bytes);
    if (!h) return nullptr;
    h->magic = OUT_MAGIC;
    h->cap = bytes;
    return h + 1;
}

void ta_free(void *p) {
    if (!p) return;
    OutHdr *h = (OutHdr *)p - 1;
    if (h->magic != OUT_MAGIC) return;  // not a live pointer this library returned (or already freed): leave it alone
    if (h->cap >= (64u << 10) && h->cap <= OUT_MAX_PARKED) {
        std::lock_guard<std::mutex> lock(g_out_mu);
        int victim = -1;
        h->magic = OUT_PARKED;
        for (int i = 0; i < OUT_SLOTS; i++) {
            if (!g_out_cache[i]) {
                g_out_cache[i] = h;
                return;
            }
            if (g_out_cache[i]->cap < h->cap && (victim < 0 || g_out_cache[i]->cap < g_out_cache[victim]->cap)) victim = i;
        }
        if (victim >= 0) {  // keep the larger block
            OutHdr *old = g_out_cache[victim];
            g_out_cache[victim] = h;
            h = old;
        }
    }
    h->magic = 0;
    free(h);
}

int ta_set_length_hint(ta_ctx *ctx, int ragged) {
    if (!ctx) return TA_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->len_hint = ragged < 0 ? -1 : (ragged ? 1 : 0);
    if (ctx->multi)
        for (int r = 0; r < ta_device_count(ctx); r++) ta_multi_sub(ctx, r)->len_hint = ctx->len_hint;
    return TA_OK;
}

void ta_trim(void) {  // hands the parked output blocks back to the allocator
    std::lock_guard<std::mutex> lock(g_out_mu);
    for (int i = 0; i < OUT_SLOTS; i++) {
        if (g_out_cache[i]) {
            g_out_cache[i]->magic = 0;
            free(g_out_cache[i]);
            g_out_cache[i] = nullptr;
        }
    }
}

int ta_costs_valid(ta_costs c) {  // EditCosts::new, reference src/levenshtein.rs:44-52
    if (c.mismatch == 0 || c.gap == 0) return 0;
    if (c.transpose != 0) {
        if ((c.transpose >> 1) >= c.mismatch) return 0;
        if ((c.transpose >> 1) >= c.gap) return 0;
    }
    return 1;
}

int ta_costs_valid_search(ta_costs c) {  // check_search, reference src/levenshtein.rs:67-71
    if (!ta_costs_valid(c)) return 0;
    if (c.transpose != 0) {
        const unsigned lim = (unsigned)c.start_gap + (unsigned)c.gap;
        if (lim > 255 || c.transpose > lim) return 0;  // the reference adds in u8: overflow is a panic in debug
    }
    return 1;
}

uint32_t ta_search_default_k(size_t needle_len) {  // reference src/levenshtein.rs:1873
    return (uint32_t)(needle_len >> 1) + ((uint32_t)needle_len & 1u);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// host-buffer batch plumbing
namespace {

struct BatchStats {
    uint64_t a_bytes = 0, b_bytes = 0;
    uint32_t max_len = 0;
    bool ragged = false;  // more than 1/8 of the pairs are in another 16-byte length class than the first pair
};

// validates monotone offsets, computes totals; returns TA_OK or an error
int scan_offsets(const uint64_t *a_off, const uint64_t *b_off, size_t n, bool need_equal, BatchStats &st) {
    uint64_t max_len = 0, bad = 0, mismatch = 0, other_class = 0;
    const uint64_t cls0 = n ? std::max(a_off[1] - a_off[0], b_off[1] - b_off[0]) >> 4 : 0;
    for (size_t i = 0; i < n; i++) {  // branch-free so that it vectorises
        const uint64_t la = a_off[i + 1] - a_off[i], lb = b_off[i + 1] - b_off[i];
        bad |= (uint64_t)(a_off[i + 1] < a_off[i]) | (uint64_t)(b_off[i + 1] < b_off[i]);
        max_len = std::max(max_len, std::max(la, lb));
        mismatch |= la ^ lb;
        other_class += (uint64_t)((std::max(la, lb) >> 4) != cls0);
    }
    st.ragged = other_class * 8 > n;
    if (bad) return TA_ERR_BAD_ARG;
    if (max_len > TA_MAX_STRING_LEN && !need_equal) return TA_ERR_TOO_LARGE;  // DP cells are u32; Hamming has no such limit
    st.a_bytes = a_off[n] - a_off[0];
    st.b_bytes = b_off[n] - b_off[0];
    st.max_len = (uint32_t)std::min<uint64_t>(max_len, 0xFFFFFFFFull);
    if (need_equal && mismatch != 0) return TA_ERR_LEN_MISMATCH;
    return TA_OK;
}

// Upload one CSR side: the byte range [off[0], off[n]) and the n+1 offsets (kept as-is; kernels index bytes
// relative to `base - off[0]`).  Returns device pointers such that dev_bytes + off[i] is string i.
int upload_side(ta_ctx *ctx, int slot, bool is_a, const uint8_t *bytes, const uint64_t *off, size_t n,
                const uint8_t **dev_bytes, const uint64_t **dev_off, cudaStream_t st) {
    DevBuf &db = is_a ? ctx->d_a[slot] : ctx->d_b[slot];
    DevBuf &dofs = is_a ? ctx->d_aoff[slot] : ctx->d_boff[slot];
    const uint64_t lo = off[0], hi = off[n];
    // keep the device copy congruent to the host buffer modulo 16 so aligned strings stay aligned
    const size_t skew = (size_t)(lo & 15);
    int rc;
    if ((rc = ta_dev_reserve(ctx, db, skew + (hi - lo) + 64)) != TA_OK) return rc;
    if ((rc = ta_dev_reserve(ctx, dofs, (n + 1) * sizeof(uint64_t))) != TA_OK) return rc;
    if (hi > lo) TA_CUDA(ctx, cudaMemcpyAsync((uint8_t *)db.p + skew, bytes + lo, hi - lo, cudaMemcpyHostToDevice, st));
    TA_CUDA(ctx, cudaMemcpyAsync(dofs.p, off, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    *dev_bytes = (const uint8_t *)db.p + skew - lo;  // virtual base: + off[i] lands inside the buffer
    *dev_off = (const uint64_t *)dofs.p;
    return TA_OK;
}

// gather the indices of pairs whose result is still TA_NONE (exponential-k driver)
__global__ void collect_none_kernel(const uint32_t *__restrict__ out, const uint32_t *__restrict__ idx_in, size_t n,
                                    uint32_t *__restrict__ idx_out, uint32_t *__restrict__ counter) {
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += (size_t)gridDim.x * blockDim.x) {
        const uint32_t pair = idx_in ? idx_in[w] : (uint32_t)w;
        if (out[pair] == TA_NONE) idx_out[atomicAdd(counter, 1u)] = pair;
    }
}

int check_costs(ta_costs c) { return ta_costs_valid(c) ? TA_OK : TA_ERR_BAD_COSTS; }

// exponential search on device-resident data: k = (16,) 30, 60, ... over the pairs that are still TA_NONE
// (reference src/levenshtein.rs:1445-1454, 1480-1494, 1516-1526)
int exp_rounds_dev(ta_ctx *ctx, const uint8_t *da, const uint64_t *da_off, const uint8_t *db, const uint64_t *db_off,
                   size_t n, ta_costs costs, uint32_t max_len, uint32_t *d_out, cudaStream_t st) {
    int rc;
    // The result is the exact distance whatever thresholds are tried, so the schedule is free.  With unit costs a
    // first round at the widest band the block-table kernel takes (k = 16, 15 with transpositions: 30 instructions per
    // column instead of 42 for the 32-row sliding table that k = 30 needs) answers every pair within that distance
    // 1.4x faster; the rest continue with the reference's 30, 60, 120, ... (TA_EXP_FIRST_K overrides, 30 = reference).
    static const int env_first = getenv("TA_EXP_FIRST_K") ? atoi(getenv("TA_EXP_FIRST_K")) : 0;
    const bool unit = costs.mismatch == 1 && costs.gap == 1 && costs.start_gap == 0 && costs.transpose <= 1;
    uint32_t k = env_first > 0 ? (uint32_t)env_first : (unit ? (costs.transpose ? 15u : 16u) : 30u);
    if (k > 30) k = 30;
    if ((rc = ta_launch_lev(ctx, da, da_off, db, db_off, n, nullptr, k, costs, max_len, d_out, st)) != TA_OK) return rc;
    if ((rc = ta_dev_reserve(ctx, ctx->d_work[0], n * sizeof(uint32_t))) != TA_OK) return rc;
    if ((rc = ta_dev_reserve(ctx, ctx->d_work[1], n * sizeof(uint32_t))) != TA_OK) return rc;
    uint32_t *idx[2] = {(uint32_t *)ctx->d_work[0].p, (uint32_t *)ctx->d_work[1].p};
    const uint32_t *cur = nullptr;
    size_t cur_n = n;
    int flip = 0;
    uint32_t *counter = ctx->d_flags + 1;
    for (;;) {
        TA_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(uint32_t), st));
        const unsigned blocks = (unsigned)std::min<size_t>((cur_n + 255) / 256, (size_t)ctx->sm_count * 8);
        collect_none_kernel<<<blocks, 256, 0, st>>>(d_out, cur, cur_n, idx[flip], counter);
        ctx->launches++;
        TA_CUDA(ctx, cudaGetLastError());
        TA_CUDA(ctx, cudaMemcpyAsync(ctx->h_flags + 1, counter, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        TA_CUDA(ctx, cudaStreamSynchronize(st));
        const uint32_t remaining = ctx->h_flags[1];
        if (remaining == 0) break;
        if (k > 0x7FFFFFFFu) return TA_ERR_TOO_LARGE;  // cannot happen: k exceeds every cost bound long before
        k = k < 30 ? 30 : k * 2;
        cur = idx[flip];
        cur_n = remaining;
        flip ^= 1;
        if ((rc = ta_launch_lev(ctx, da, da_off, db, db_off, cur_n, cur, k, costs, max_len, d_out, st)) != TA_OK)
            return rc;
    }
    return TA_OK;
}

enum Op { OP_HAMMING, OP_LEV_K, OP_LEV_EXP };

int run_pairs(ta_ctx *ctx, Op op, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
              size_t n, uint32_t k, ta_costs costs, uint32_t *out) {
    if (!ctx) return TA_ERR_BAD_ARG;
    if (n == 0) return TA_OK;
    if (!a_off || !b_off || !out) return TA_ERR_BAD_ARG;
    if (n > 0xFFFFFFF0ull) return TA_ERR_TOO_LARGE;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (ctx->multi) {
        // one call, several GPUs: contiguous ranges of pairs balanced by bytes, each through the
        // single-device path of its own sub-context and host thread, results straight into the caller's slices
        const uint64_t total = (a_off[n] - a_off[0]) + (b_off[n] - b_off[0]);
        const int parts = ta_multi_parts(ctx, total, n);
        std::vector<size_t> bound;
        ta_multi_bounds(a_off, b_off, n, parts, bound);
        return ta_multi_run(ctx, parts, [&](int r) -> int {
            const size_t lo = bound[r], cnt = bound[r + 1] - bound[r];
            if (cnt == 0) return TA_OK;
            return run_pairs(ta_multi_sub(ctx, r), op, a, a_off + lo, b, b_off + lo, cnt, k, costs, out + lo);
        });
    }
    TA_CUDA(ctx, cudaSetDevice(ctx->device));
    // Pipeline: the batch is cut into up to MAX_CHUNKS ranges of pairs.  All H2D copies are queued on the copy stream
    // first (a range only needs its first and last offsets), the offsets are validated on the host while the DMA
    // engine is busy, then each range's kernel + D2H is queued on the compute stream behind that range's copy event,
    // so kernels overlap the copies of later ranges.  A contract violation found by the validation drains the
    // streams and returns the error.
    if (a_off[n] < a_off[0] || b_off[n] < b_off[0]) return TA_ERR_BAD_ARG;
    if ((a_off[n] > a_off[0] && !a) || (b_off[n] > b_off[0] && !b)) return TA_ERR_BAD_ARG;
    int rc;
    cudaStream_t st = ctx->stream, cp = ctx->stream2;
    const uint64_t a_lo = a_off[0], b_lo = b_off[0];
    const uint64_t a_bytes = a_off[n] - a_lo, b_bytes = b_off[n] - b_lo;
    const size_t a_skew = (size_t)(a_lo & 15), b_skew = (size_t)(b_lo & 15);  // keep 16-byte alignment classes
    if ((rc = ta_dev_reserve(ctx, ctx->d_a[0], a_skew + a_bytes + 64)) != TA_OK) return rc;
    if ((rc = ta_dev_reserve(ctx, ctx->d_b[0], b_skew + b_bytes + 64)) != TA_OK) return rc;
    if ((rc = ta_dev_reserve(ctx, ctx->d_aoff[0], (n + 1) * sizeof(uint64_t))) != TA_OK) return rc;
    if ((rc = ta_dev_reserve(ctx, ctx->d_boff[0], (n + 1) * sizeof(uint64_t))) != TA_OK) return rc;
    if ((rc = ta_dev_reserve(ctx, ctx->d_out[0], n * sizeof(uint32_t))) != TA_OK) return rc;
    uint8_t *buf_a = (uint8_t *)ctx->d_a[0].p + a_skew, *buf_b = (uint8_t *)ctx->d_b[0].p + b_skew;
    const uint8_t *da = buf_a - a_lo, *db = buf_b - b_lo;  // virtual bases: + off[i] lands inside the buffers
    uint64_t *da_off = (uint64_t *)ctx->d_aoff[0].p, *db_off = (uint64_t *)ctx->d_boff[0].p;
    uint32_t *d_out = (uint32_t *)ctx->d_out[0].p;

    const int chunks = (int)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)ta_ctx::MAX_CHUNKS, (uint64_t)n,
                                                                      (a_bytes + b_bytes) >> 24}));  // >= 16 MB each
    size_t bound[ta_ctx::MAX_CHUNKS + 1];
    for (int c = 0; c <= chunks; c++) bound[c] = n * (size_t)c / chunks;
    TA_CUDA(ctx, cudaEventRecord(ctx->ev_h2d[0], st));  // the copy stream must not overwrite buffers still in use
    TA_CUDA(ctx, cudaStreamWaitEvent(cp, ctx->ev_h2d[0], 0));
    for (int c = 0; c < chunks; c++) {
        const size_t lo = bound[c], hi = bound[c + 1];
        if (a_off[hi] < a_off[lo] || b_off[hi] < b_off[lo] || a_off[lo] < a_lo || b_off[lo] < b_lo ||
            a_off[hi] > a_off[n] || b_off[hi] > b_off[n]) {
            cudaStreamSynchronize(cp);
            return TA_ERR_BAD_ARG;
        }
        if (a_off[hi] > a_off[lo])
            TA_CUDA(ctx, cudaMemcpyAsync(buf_a + (a_off[lo] - a_lo), a + a_off[lo], a_off[hi] - a_off[lo],
                                         cudaMemcpyHostToDevice, cp));
        if (b_off[hi] > b_off[lo])
            TA_CUDA(ctx, cudaMemcpyAsync(buf_b + (b_off[lo] - b_lo), b + b_off[lo], b_off[hi] - b_off[lo],
                                         cudaMemcpyHostToDevice, cp));
        TA_CUDA(ctx, cudaMemcpyAsync(da_off + lo, a_off + lo, (hi - lo + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, cp));
        TA_CUDA(ctx, cudaMemcpyAsync(db_off + lo, b_off + lo, (hi - lo + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, cp));
        TA_CUDA(ctx, cudaEventRecord(ctx->ev_h2d[c], cp));
    }
    BatchStats bs;
    rc = scan_offsets(a_off, b_off, n, op == OP_HAMMING, bs);
    if (rc != TA_OK) {
        cudaStreamSynchronize(cp);
        return rc;
    }
    ctx->batch_ragged = ctx->len_hint < 0 ? bs.ragged : ctx->len_hint != 0;
    if (op == OP_HAMMING) TA_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, sizeof(uint32_t), st));
    for (int c = 0; c < chunks && rc == TA_OK; c++) {
        const size_t lo = bound[c], cn = bound[c + 1] - bound[c];
        TA_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_h2d[c], 0));
        if (cn == 0) continue;
        switch (op) {
            case OP_HAMMING: {
                const uint32_t avg = (uint32_t)std::min<uint64_t>(bs.a_bytes / n, 0xFFFFFFFFull);
                rc = ta_launch_hamming(ctx, da, da_off + lo, db, db_off + lo, cn, avg, d_out + lo, ctx->d_flags, st);
                break;
            }
            case OP_LEV_K:
                rc = ta_launch_lev(ctx, da, da_off + lo, db, db_off + lo, cn, nullptr, k, costs, bs.max_len, d_out + lo, st);
                break;
            case OP_LEV_EXP:
                rc = exp_rounds_dev(ctx, da, da_off + lo, db, db_off + lo, cn, costs, bs.max_len, d_out + lo, st);
                break;
        }
        if (rc == TA_OK)
            TA_CUDA(ctx, cudaMemcpyAsync(out + lo, d_out + lo, cn * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    if (rc != TA_OK) {
        cudaStreamSynchronize(cp);
        cudaStreamSynchronize(st);
        return rc;
    }
    TA_CUDA(ctx, cudaStreamSynchronize(st));
    return TA_OK;
}

}  // namespace

extern "C" {

int ta_hamming_batch(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                     size_t n, uint32_t *out) {
    ta_costs none = {1, 1, 0, 0};
    return run_pairs(ctx, OP_HAMMING, a, a_off, b, b_off, n, 0, none, out);
}

int ta_levenshtein_k_batch(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                           const uint64_t *b_off, size_t n, uint32_t k, ta_costs costs, uint32_t *out) {
    int rc = check_costs(costs);
    if (rc != TA_OK) return rc;
    return run_pairs(ctx, OP_LEV_K, a, a_off, b, b_off, n, k, costs, out);
}

int ta_levenshtein_exp_batch(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                             const uint64_t *b_off, size_t n, ta_costs costs, uint32_t *out) {
    int rc = check_costs(costs);
    if (rc != TA_OK) return rc;
    return run_pairs(ctx, OP_LEV_EXP, a, a_off, b, b_off, n, 0, costs, out);
}

}  // extern "C"

namespace {

// Runs the TRACE kernel + the two walk passes over `cnt` work items (pairs pair_base + w, or h_idx[w] when an index
// list is given; d_idx is its device copy) in chunks bounded by the trace workspace, appending every pair's runs
// to `pool` and recording where they start (pos) and how many there are (num).
int run_trace_group(ta_ctx *ctx, cudaStream_t st, const uint8_t *da, const uint64_t *da_off, const uint8_t *db,
                    const uint64_t *db_off, const uint32_t *d_idx, const uint32_t *h_idx, size_t pair_base, size_t cnt,
                    uint32_t k, ta_costs costs, uint32_t max_len, uint32_t *d_out, std::vector<ta_edit> &pool,
                    std::vector<uint64_t> &pos, std::vector<uint32_t> &num) {
    if (cnt == 0) return TA_OK;
    const uint32_t W = ta_band_width_bound(k, costs, max_len);
    const uint32_t wc = ta_trace_cells(W);
    if (wc == 0) return TA_ERR_TOO_LARGE;  // traceback supports bands of up to 1024 diagonals
    const size_t stride = (((size_t)2 * max_len + 3) * wc + 15) & ~(size_t)15;
    const size_t budget = (size_t)2 << 30;
    const size_t chunk_n = std::max<size_t>(1, std::min(cnt, budget / stride));
    int rc;
    if ((rc = ta_dev_reserve(ctx, ctx->d_work[3], chunk_n * stride)) != TA_OK) return rc;
    if ((rc = ta_dev_reserve(ctx, ctx->d_work[2], chunk_n * sizeof(uint32_t))) != TA_OK) return rc;
    if ((rc = ta_dev_reserve(ctx, ctx->d_work[1], chunk_n * sizeof(uint64_t))) != TA_OK) return rc;
    std::vector<uint32_t> h_counts(chunk_n);
    std::vector<uint64_t> h_off(chunk_n + 1);
    for (size_t c0 = 0; c0 < cnt; c0 += chunk_n) {
        const size_t cn = std::min(chunk_n, cnt - c0);
        const uint32_t *ci = d_idx ? d_idx + c0 : nullptr;
        uint8_t *trace = (uint8_t *)ctx->d_work[3].p;
        uint32_t *d_counts = (uint32_t *)ctx->d_work[2].p;
        uint64_t *d_off = (uint64_t *)ctx->d_work[1].p;
        if ((rc = ta_launch_lev_band_trace(ctx, da, da_off, db, db_off, cn, ci, pair_base + c0, k, costs, max_len, d_out,
                                           trace, stride, st)) != TA_OK)
            return rc;
        if ((rc = ta_launch_trace_walk(ctx, da, da_off, db, db_off, cn, ci, pair_base + c0, k, costs, max_len, d_out,
                                       trace, stride, d_counts, nullptr, nullptr, st)) != TA_OK)
            return rc;
        TA_CUDA(ctx, cudaMemcpyAsync(h_counts.data(), d_counts, cn * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        TA_CUDA(ctx, cudaStreamSynchronize(st));
        h_off[0] = 0;
        for (size_t q = 0; q < cn; q++) h_off[q + 1] = h_off[q] + h_counts[q];
        const uint64_t total = h_off[cn];
        if (total) {
            if ((rc = ta_dev_reserve(ctx, ctx->d_work[0], total * sizeof(ta_edit))) != TA_OK) return rc;
            TA_CUDA(ctx, cudaMemcpyAsync(d_off, h_off.data(), cn * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
            if ((rc = ta_launch_trace_walk(ctx, da, da_off, db, db_off, cn, ci, pair_base + c0, k, costs, max_len,
                                           d_out, trace, stride, d_counts, d_off, (ta_edit *)ctx->d_work[0].p, st)) != TA_OK)
                return rc;
            const size_t base = pool.size();
            pool.resize(base + total);
            TA_CUDA(ctx, cudaMemcpyAsync(pool.data() + base, ctx->d_work[0].p, total * sizeof(ta_edit),
                                         cudaMemcpyDeviceToHost, st));
            TA_CUDA(ctx, cudaStreamSynchronize(st));
            for (size_t q = 0; q < cn; q++) {
                const size_t pair = h_idx ? h_idx[c0 + q] : pair_base + c0 + q;
                pos[pair] = base + h_off[q];
                num[pair] = h_counts[q];
            }
        }
    }
    return TA_OK;
}

int trace_batch(ta_ctx *ctx, bool exp_mode, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                const uint64_t *b_off, size_t n, uint32_t k, ta_costs costs, uint32_t *out_dist, ta_edit **out_edits,
                uint64_t **out_edit_off) {
    if (!ctx || !out_edits || !out_edit_off) return TA_ERR_BAD_ARG;
    *out_edits = nullptr;
    *out_edit_off = nullptr;
    int rc = check_costs(costs);
    if (rc != TA_OK) return rc;
    if (n && (!a_off || !b_off || !out_dist)) return TA_ERR_BAD_ARG;
    if (n > 0xFFFFFFF0ull) return TA_ERR_TOO_LARGE;
    if (ctx->multi && n) {
        std::lock_guard<std::mutex> lock(ctx->mu);
        const uint64_t total = (a_off[n] - a_off[0]) + (b_off[n] - b_off[0]);
        const int parts = ta_multi_parts(ctx, total, n);
        std::vector<size_t> bound;
        ta_multi_bounds(a_off, b_off, n, parts, bound);
        std::vector<ta_edit *> ed(parts, nullptr);
        std::vector<uint64_t *> eo(parts, nullptr);
        rc = ta_multi_run(ctx, parts, [&](int r) -> int {
            const size_t lo = bound[r], cnt = bound[r + 1] - bound[r];
            if (cnt == 0) return TA_OK;
            return trace_batch(ta_multi_sub(ctx, r), exp_mode, a, a_off + lo, b, b_off + lo, cnt, k, costs,
                               out_dist + lo, &ed[r], &eo[r]);
        });
        if (rc != TA_OK) {
            for (int r = 0; r < parts; r++) ta_free(ed[r]), ta_free(eo[r]);
            return rc;
        }
        return ta_concat_lists<ta_edit>(parts, bound, n, ed, eo, out_edits, out_edit_off);
    }
    uint64_t *eoff = (uint64_t *)ta_out_alloc((n + 1) * sizeof(uint64_t));
    if (eoff) memset(eoff, 0, (n + 1) * sizeof(uint64_t));
    if (!eoff) return TA_ERR_NOMEM;
    std::vector<ta_edit> pool;
    std::vector<uint64_t> pos(n, 0);
    std::vector<uint32_t> num(n, 0);
    if (n) {
        std::lock_guard<std::mutex> lock(ctx->mu);
        auto run = [&]() -> int {
            TA_CUDA(ctx, cudaSetDevice(ctx->device));
            BatchStats bs;
            int r = scan_offsets(a_off, b_off, n, false, bs);
            ctx->batch_ragged = ctx->len_hint < 0 ? bs.ragged : ctx->len_hint != 0;
            if (r != TA_OK) return r;
            if ((bs.a_bytes && !a) || (bs.b_bytes && !b)) return TA_ERR_BAD_ARG;
            cudaStream_t st = ctx->stream;
            const uint8_t *da, *db;
            const uint64_t *da_off, *db_off;
            if ((r = upload_side(ctx, 0, true, a, a_off, n, &da, &da_off, st)) != TA_OK) return r;
            if ((r = upload_side(ctx, 0, false, b, b_off, n, &db, &db_off, st)) != TA_OK) return r;
            if ((r = ta_dev_reserve(ctx, ctx->d_out[0], n * sizeof(uint32_t))) != TA_OK) return r;
            uint32_t *d_out = (uint32_t *)ctx->d_out[0].p;
            // Distances first -- the k-bounded call (src/levenshtein.rs:714-827) or the exponential search
            // (:1486-1493); then the traceback of every pair that has a distance, in groups by distance with the
            // thresholds 30, 60, 120, ...: the decisions along an optimal path do not depend on the band width once
            // the band holds the path, so a pair at distance d is traced in the band of min(k, first threshold >= d)
            // whatever k the caller passed (k = u32::MAX on long strings must not ask for a band as wide as the
            // strings).  Only a pair whose OWN distance needs more than 1024 diagonals is TA_ERR_TOO_LARGE.
            if (exp_mode)
                r = exp_rounds_dev(ctx, da, da_off, db, db_off, n, costs, bs.max_len, d_out, st);
            else
                r = ta_launch_lev(ctx, da, da_off, db, db_off, n, nullptr, k, costs, bs.max_len, d_out, st);
            if (r != TA_OK) return r;
            TA_CUDA(ctx, cudaMemcpyAsync(out_dist, d_out, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            TA_CUDA(ctx, cudaStreamSynchronize(st));
            std::vector<std::vector<uint32_t>> rounds;
            for (size_t i = 0; i < n; i++) {
                if (out_dist[i] == TA_NONE) continue;
                size_t rr = 0;
                uint64_t kk = 30;
                while (kk < out_dist[i]) kk *= 2, rr++;
                if (rounds.size() <= rr) rounds.resize(rr + 1);
                rounds[rr].push_back((uint32_t)i);
            }
            if ((r = ta_dev_reserve(ctx, ctx->d_aoff[1], n * sizeof(uint32_t))) != TA_OK) return r;
            uint32_t *d_idx = (uint32_t *)ctx->d_aoff[1].p;
            uint64_t kk = 30;
            for (size_t rr = 0; rr < rounds.size(); rr++, kk *= 2) {
                if (rounds[rr].empty()) continue;
                TA_CUDA(ctx, cudaMemcpyAsync(d_idx, rounds[rr].data(), rounds[rr].size() * sizeof(uint32_t),
                                             cudaMemcpyHostToDevice, st));
                const uint64_t kr = exp_mode ? kk : std::min<uint64_t>(kk, k);
                r = run_trace_group(ctx, st, da, da_off, db, db_off, d_idx, rounds[rr].data(), 0, rounds[rr].size(),
                                    (uint32_t)std::min<uint64_t>(kr, 0xFFFFFFFFull), costs, bs.max_len, d_out, pool, pos,
                                    num);
                if (r != TA_OK) return r;
            }
            TA_CUDA(ctx, cudaMemcpyAsync(out_dist, d_out, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            TA_CUDA(ctx, cudaStreamSynchronize(st));
            return TA_OK;
        };
        rc = run();
        if (rc != TA_OK) {
            cudaStreamSynchronize(ctx->stream);
            ta_free(eoff);
            return rc;
        }
    }
    for (size_t i = 0; i < n; i++) eoff[i + 1] = eoff[i] + num[i];
    ta_edit *ed = (ta_edit *)ta_out_alloc((eoff[n] ? eoff[n] : 1) * sizeof(ta_edit));
    if (!ed) {
        ta_free(eoff);
        return TA_ERR_NOMEM;
    }
    for (size_t i = 0; i < n; i++)
        if (num[i]) memcpy(ed + eoff[i], pool.data() + pos[i], (size_t)num[i] * sizeof(ta_edit));
    *out_edits = ed;
    *out_edit_off = eoff;
    return TA_OK;
}

}  // namespace

extern "C" {

int ta_levenshtein_k_trace_batch(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                                 const uint64_t *b_off, size_t n, uint32_t k, ta_costs costs, uint32_t *out_dist,
                                 ta_edit **out_edits, uint64_t **out_edit_off) {
    return trace_batch(ctx, false, a, a_off, b, b_off, n, k, costs, out_dist, out_edits, out_edit_off);
}

int ta_levenshtein_exp_trace_batch(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                                   const uint64_t *b_off, size_t n, ta_costs costs, uint32_t *out_dist,
                                   ta_edit **out_edits, uint64_t **out_edit_off) {
    return trace_batch(ctx, true, a, a_off, b, b_off, n, 0, costs, out_dist, out_edits, out_edit_off);
}

int ta_hamming_batch_dev(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                         const uint64_t *b_off, size_t n, uint32_t *out, void *stream) {
    if (!ctx || ctx->multi) return TA_ERR_BAD_ARG;  // device pointers belong to one device: use a single-device ctx
    if (n == 0) return TA_OK;
    if (!a_off || !b_off || !out) return TA_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    TA_CUDA(ctx, cudaSetDevice(ctx->device));
    // mean length is unknown without reading device offsets; 64-byte strings and up use >= 4 lanes per pair
    // (ta_hamming_batch_dev_len takes it from the caller)
    return ta_launch_hamming(ctx, a, a_off, b, b_off, n, 64, out, ctx->d_flags, (cudaStream_t)stream);
}

int ta_hamming_batch_dev_len(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                             const uint64_t *b_off, size_t n, uint32_t mean_len, uint32_t *out, void *stream) {
    if (!ctx || ctx->multi) return TA_ERR_BAD_ARG;
    if (n == 0) return TA_OK;
    if (!a_off || !b_off || !out) return TA_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    TA_CUDA(ctx, cudaSetDevice(ctx->device));
    return ta_launch_hamming(ctx, a, a_off, b, b_off, n, mean_len ? mean_len : 64, out, ctx->d_flags, (cudaStream_t)stream);
}

int ta_levenshtein_k_batch_dev(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                               const uint64_t *b_off, size_t n, uint32_t k, ta_costs costs, uint32_t max_len,
                               uint32_t *out, void *stream) {
    if (!ctx || ctx->multi) return TA_ERR_BAD_ARG;
    int rc = check_costs(costs);
    if (rc != TA_OK) return rc;
    if (n == 0) return TA_OK;
    if (!a_off || !b_off || !out) return TA_ERR_BAD_ARG;
    if (max_len > TA_MAX_STRING_LEN) return TA_ERR_TOO_LARGE;
    std::lock_guard<std::mutex> lock(ctx->mu);
    TA_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->batch_ragged = ctx->len_hint > 0;
    return ta_launch_lev(ctx, a, a_off, b, b_off, n, nullptr, k, costs, max_len, out, (cudaStream_t)stream);
}

int ta_levenshtein_exp_batch_dev(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                                 const uint64_t *b_off, size_t n, ta_costs costs, uint32_t max_len, uint32_t *out,
                                 void *stream) {
    if (!ctx || ctx->multi) return TA_ERR_BAD_ARG;
    int rc = check_costs(costs);
    if (rc != TA_OK) return rc;
    if (n == 0) return TA_OK;
    if (!a_off || !b_off || !out) return TA_ERR_BAD_ARG;
    if (max_len > TA_MAX_STRING_LEN || n > 0xFFFFFFF0ull) return TA_ERR_TOO_LARGE;
    std::lock_guard<std::mutex> lock(ctx->mu);
    TA_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->batch_ragged = ctx->len_hint > 0;
    return exp_rounds_dev(ctx, a, a_off, b, b_off, n, costs, max_len, out, (cudaStream_t)stream);
}

int ta_dev_status(ta_ctx *ctx, void *stream) {
    if (!ctx || ctx->multi) return TA_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    TA_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    TA_CUDA(ctx, cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    TA_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, sizeof(uint32_t), st));
    TA_CUDA(ctx, cudaStreamSynchronize(st));
    const uint32_t f = ctx->h_flags[0];
    return f ? -(int)f : TA_OK;
}

int ta_hamming(ta_ctx *ctx, const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len, uint32_t *out) {
    const uint64_t ao[2] = {0, a_len}, bo[2] = {0, b_len};
    return ta_hamming_batch(ctx, a, ao, b, bo, 1, out);
}

int ta_levenshtein_simd_k_with_opts(ta_ctx *ctx, const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len,
                                    uint32_t k, ta_costs costs, uint32_t *out) {
    const uint64_t ao[2] = {0, a_len}, bo[2] = {0, b_len};
    return ta_levenshtein_k_batch(ctx, a, ao, b, bo, 1, k, costs, out);
}

int ta_levenshtein_exp_with_opts(ta_ctx *ctx, const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len,
                                 ta_costs costs, uint32_t *out) {
    const uint64_t ao[2] = {0, a_len}, bo[2] = {0, b_len};
    return ta_levenshtein_exp_batch(ctx, a, ao, b, bo, 1, costs, out);
}

// search entry point lives in search.cu

}  // extern "C"
