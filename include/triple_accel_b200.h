/*
 * triple_accel_b200.h -- C ABI of libtriple_accel_b200.so
 *
 * A B200 (sm_100a) batched edit-distance engine that keeps the contracts of the Rust crate `triple_accel`
 * (reference @ 0f2119a, v0.4.0).  The reference has no FFI: its boundary is the crate's public Rust functions
 * (re-exported at src/lib.rs:126-127).  Each entry point below names the reference function whose contract it
 * implements; INTEGRATION.md shows the Rust shim (`extern "C"` block + safe wrappers with the crate's names) a
 * maintainer would add.  Plain pointers and sizes only; nothing is retained after a call returns.
 *
 * Batch layout ("CSR"): a batch of n byte strings is one contiguous byte buffer plus n+1 u64 offsets;
 * string i is bytes[off[i] .. off[i+1]).  Pair i is (a_i, b_i).
 *
 * Threading: a ta_ctx owns one CUDA device (ta_init) or several (ta_init_multi), their streams and staging buffers;
 * calls on one ctx are serialised by an internal mutex, any number of threads may share a ctx, and distinct contexts
 * run concurrently (also on the same device).  The reference's functions are pure and re-entrant; so are these.
 *
 * Errors: 0 = success; negative = error (ta_strerror).  Contract violations that the reference turns into
 * panics are reported as error codes (the Rust shim panics on them):  TA_ERR_LEN_MISMATCH <-> assert at
 * src/hamming.rs:38,318;  TA_ERR_BAD_COSTS <-> asserts at src/levenshtein.rs:44-52,69.
 * "Not within k" is a value (TA_NONE), not an error (src/levenshtein.rs:428-430, 539-541, 860-862, 1166-1168).
 * There is no CPU fallback: without a usable CUDA device every compute entry point fails with TA_ERR_CUDA.
 */
#ifndef TRIPLE_ACCEL_B200_H
#define TRIPLE_ACCEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TA_ABI_VERSION 2

/* Option::None for a distance (Option<u32> in src/levenshtein.rs:342, 677, 714-720) */
#define TA_NONE 0xFFFFFFFFu

enum {
    TA_OK = 0,
    TA_ERR_CUDA = -1,         /* CUDA runtime failure (ta_last_error has the text) */
    TA_ERR_LEN_MISMATCH = -2, /* hamming: a_i and b_i differ in length (src/hamming.rs:38, 318) */
    TA_ERR_BAD_COSTS = -3,    /* EditCosts::new / check_search asserts (src/levenshtein.rs:44-52, 69) */
    TA_ERR_BAD_ARG = -4,      /* null pointer, non-monotone offsets, bad enum value */
    TA_ERR_TOO_LARGE = -5,    /* a string exceeds TA_MAX_STRING_LEN, or cost arithmetic would overflow u32 */
    TA_ERR_NOMEM = -6,        /* host allocation failed */
    TA_ERR_NUL_BYTE = -7      /* hamming_search: NUL byte in a haystack (check_no_null_bytes, src/lib.rs:237-243) */
};

/* (len_a + len_b) * max(cost) must stay below 2^30 so that u32 cells never wrap */
#define TA_MAX_STRING_LEN (1u << 21)

/* EditCosts (src/levenshtein.rs:20-26).  transpose == 0 encodes transpose_cost: None
 * (Some(t) requires t > 0, src/levenshtein.rs:47-48). */
typedef struct {
    uint8_t mismatch, gap, start_gap, transpose;
} ta_costs;

/* LEVENSHTEIN_COSTS (src/levenshtein.rs:76-81), RDAMERAU_COSTS (src/levenshtein.rs:84-89) */
#define TA_LEVENSHTEIN_COSTS_INIT {1, 1, 0, 0}
#define TA_RDAMERAU_COSTS_INIT {1, 1, 0, 1}

/* Match (src/lib.rs:134-142): start inclusive, end exclusive, k = cost of the match */
typedef struct {
    uint64_t start, end;
    uint32_t k, _pad;
} ta_match;

/* Edit (src/lib.rs:159-165): a run of `count` edits of one EditType (src/lib.rs:147-154) */
enum { TA_EDIT_MATCH = 0, TA_EDIT_MISMATCH = 1, TA_EDIT_AGAP = 2, TA_EDIT_BGAP = 3, TA_EDIT_TRANSPOSE = 4 };
typedef struct {
    uint32_t edit, count;
} ta_edit;

/* SearchType (src/lib.rs:170-174) */
enum { TA_SEARCH_ALL = 0, TA_SEARCH_BEST = 1 };

typedef struct ta_ctx ta_ctx;

/* ---- lifetime -------------------------------------------------------------------------------------------- */
int ta_abi_version(void);
/* Create a context on CUDA device `device` (one context per GPU; one process per GPU in multi-GPU runs). */
int ta_init(int device, ta_ctx **out);
/* Create ONE context over `n_devices` distinct CUDA devices of this box.  Every host-buffer batch entry
 * point below then splits its batch into contiguous ranges of pairs / haystacks balanced by bytes, runs each range on
 * its own device (own streams, staging buffers and host thread) and writes the results into the caller's arrays in
 * batch order: same results as a single-device context, no data-path collective.  The needle of
 * ta_levenshtein_search_batch is uploaded to the first device and sent to the others with ncclBroadcast over a
 * single-process communicator (libnccl.so.2 is loaded at run time; without it the needle is copied per device, see
 * ta_multi_uses_nccl).  Batches below ~4 MB per device use fewer devices.  The *_dev entry points take pointers that
 * live on one device and therefore return TA_ERR_BAD_ARG on a multi-device context.  n_devices == 1 is ta_init. */
int ta_init_multi(const int *devices, int n_devices, ta_ctx **out);
int ta_device_count(ta_ctx *ctx);               /* devices this ctx spans */
/* The split itself, as pure host arithmetic (no device needed): bounds_out[0 .. parts] with units bounds_out[r] ..
 * bounds_out[r + 1] going to device r -- contiguous ranges with equal shares of the bytes (b_off may be NULL). */
int ta_shard_bounds(const uint64_t *a_off, const uint64_t *b_off, size_t n, int parts, uint64_t *bounds_out);
int ta_multi_uses_nccl(ta_ctx *ctx);            /* 1 = needle travels by ncclBroadcast */
uint64_t ta_multi_needle_broadcasts(ta_ctx *ctx); /* ncclBroadcast calls issued so far */
void ta_shutdown(ta_ctx *ctx);
const char *ta_strerror(int code);
const char *ta_last_error(ta_ctx *ctx); /* text of the last CUDA error seen by this ctx */
int ta_device(ta_ctx *ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t ta_launch_count(ta_ctx *ctx);

/* Length hint for the unit-cost distance kernels of this ctx.  Batches whose pairs differ in length by more than a
 * 16-byte class (say 96..160 bytes) run ~20 % faster when the kernel orders each tile of pairs by length first
 * (lev_bitpar_duo_tiled_kernel); equal-length batches are ~5 % faster without that pass.  The host-buffer entry
 * points see the offsets and decide per batch; the *_dev entry points cannot (the offsets are on the device), so the
 * caller may say: ragged = 1 (lengths vary), 0 (equal lengths), -1 (default: host-buffer calls decide from the
 * offsets, *_dev calls assume equal lengths).  Results never depend on the hint.  On a multi-device ctx it applies
 * to every device. */
int ta_set_length_hint(ta_ctx *ctx, int ragged);

/* Pinned host memory for fast H2D/D2H of batch buffers (optional: any host pointer is accepted). */
void *ta_host_alloc(size_t bytes);
void ta_host_free(void *p);
/* Releases an output array this library returned (match lists, per-unit offsets, edit lists).  Pass only such
 * pointers (NULL is ignored): the blocks carry a small header and large ones are parked for reuse by the next call
 * instead of going back to the system allocator. */
void ta_free(void *p);
/* Returns the parked blocks of ta_free to the system allocator (at most six blocks of <= 64 MB are ever parked). */
void ta_trim(void);

/* EditCosts::new validity (src/levenshtein.rs:38-60) / check_search (src/levenshtein.rs:67-71): 1 = valid */
int ta_costs_valid(ta_costs c);
int ta_costs_valid_search(ta_costs c);

/* ---- host-buffer batch entry points (H2D + kernels + D2H inside the call) ---------------------------------- */

/* hamming (src/hamming.rs:390-392) == hamming_naive (src/hamming.rs:36-47) per pair.
 * out[i] = number of positions where a_i and b_i differ.  TA_ERR_LEN_MISMATCH if any pair differs in length. */
int ta_hamming_batch(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b, const uint64_t *b_off,
                     size_t n, uint32_t *out);

/* levenshtein_simd_k_with_opts(a, b, k, false, costs) (src/levenshtein.rs:714-827), bit-exact with the scalar
 * levenshtein_naive_k_with_opts (src/levenshtein.rs:376-545): out[i] = d if d <= k else TA_NONE.
 * k = 0xFFFFFFFF gives levenshtein() / rdamerau() (src/levenshtein.rs:1397-1399, 1419-1423). */
int ta_levenshtein_k_batch(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                           const uint64_t *b_off, size_t n, uint32_t k, ta_costs costs, uint32_t *out);

/* levenshtein_simd_k_with_opts(a, b, k, true, costs): distances as above plus the edit traceback of every pair
 * within k, bit-exact with the scalar levenshtein_naive_k_with_opts(.., trace_on = true)
 * (src/levenshtein.rs:493-606).  Runs of pair i are (*out_edits)[(*out_edit_off)[i] .. (*out_edit_off)[i+1])
 * (none for TA_NONE pairs); release both arrays with ta_free.  Bands up to 1024 diagonals. */
int ta_levenshtein_k_trace_batch(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                                 const uint64_t *b_off, size_t n, uint32_t k, ta_costs costs, uint32_t *out_dist,
                                 ta_edit **out_edits, uint64_t **out_edit_off);
/* levenshtein_exp_with_opts(a, b, true, costs) (src/levenshtein.rs:1480-1494): exact distance and traceback. */
int ta_levenshtein_exp_trace_batch(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                                   const uint64_t *b_off, size_t n, ta_costs costs, uint32_t *out_dist,
                                   ta_edit **out_edits, uint64_t **out_edit_off);

/* levenshtein_exp / levenshtein_exp_with_opts / rdamerau_exp (src/levenshtein.rs:1445-1454, 1480-1494,
 * 1516-1526): exact distance by running the k-bounded routine with k = 30, 60, 120, ... on the pairs that are
 * still TA_NONE.  Never returns TA_NONE. */
int ta_levenshtein_exp_batch(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                             const uint64_t *b_off, size_t n, ta_costs costs, uint32_t *out);

/* levenshtein_search_simd_with_opts(needle, haystack_i, k, search_type, costs, anchored)
 * (src/levenshtein.rs:1911-2155), bit-exact with levenshtein_search_naive_with_opts (src/levenshtein.rs:1589-1838)
 * for every haystack of the batch.  *out_matches receives all matches, haystack by haystack, in the order the
 * reference iterator yields them; matches of haystack i are (*out_matches)[(*out_match_off)[i] ..
 * (*out_match_off)[i+1]).  Both arrays are allocated by the library: release with ta_free.  The needle is limited
 * to TA_MAX_STRING_LEN bytes (needles up to 256 bytes run on the warp kernel, up to ~450 bytes -- ~300 with
 * transpositions -- with the DP rows in shared memory, longer ones with the rows in a device workspace of
 * 16 (24) bytes per needle byte and resident thread), a haystack to 2^32 - 16 bytes (TA_ERR_TOO_LARGE beyond).
 * levenshtein_search(needle, haystack) (src/levenshtein.rs:2508-2513) is k = ta_search_default_k(needle_len),
 * TA_SEARCH_BEST, unit costs, anchored = 0. */
int ta_levenshtein_search_batch(ta_ctx *ctx, const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                                const uint64_t *hay_off, size_t n, uint32_t k, int search_type, ta_costs costs,
                                int anchored, ta_match **out_matches, uint64_t **out_match_off);
uint32_t ta_search_default_k(size_t needle_len); /* src/levenshtein.rs:1873 */

/* hamming_search_simd_with_opts(needle, haystack_i, k, search_type) (src/hamming.rs:454-475), bit-exact with the
 * scalar hamming_search_naive_with_opts (src/hamming.rs:96-146) for every haystack of the batch; output layout as
 * ta_levenshtein_search_batch.  A NUL byte in a searched haystack is TA_ERR_NUL_BYTE (the reference panics).
 * hamming_search(needle, haystack) (src/hamming.rs:588-590) is k = ta_search_default_k(needle_len), TA_SEARCH_BEST. */
int ta_hamming_search_batch(ta_ctx *ctx, const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                            const uint64_t *hay_off, size_t n, uint32_t k, int search_type, ta_match **out_matches,
                            uint64_t **out_match_off);

/* hamming_search_naive_with_opts (src/hamming.rs:96-146): as ta_hamming_search_batch, but NUL bytes in a haystack
 * are ordinary bytes (the scalar routine has no such restriction; only the SIMD entry checks, src/hamming.rs:463). */
int ta_hamming_search_naive_batch(ta_ctx *ctx, const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                                  const uint64_t *hay_off, size_t n, uint32_t k, int search_type,
                                  ta_match **out_matches, uint64_t **out_match_off);

/* ---- device-resident entry points (kernel-only; all pointers are device pointers on ctx's device) ---------- */
/* `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Calls are asynchronous.
 * `max_len` is an upper bound on the length of any string in the batch (picks the kernel variant).
 * Ordering: the kernels use scratch buffers that belong to the context (counters, index lists, wide-band and
 * search workspaces), so all work submitted through ONE context must be ordered by the caller -- use one stream per
 * context, or make a later call's stream wait on the earlier one; two streams that should overlap need two contexts
 * (contexts are cheap: a few streams, events and lazily grown buffers).
 * Buffers: the string buffers must be readable from the 16-byte boundary at or below their first byte to the
 * 16-byte boundary at or above their last one (the general-cost kernel stages whole aligned 16-byte vectors);
 * anything returned by cudaMalloc / a caching allocator satisfies this, a tight sub-allocation inside a larger
 * buffer does too as long as it does not end in the buffer's last 15 bytes. */
int ta_hamming_batch_dev(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                         const uint64_t *b_off, size_t n, uint32_t *out, void *stream);
/* The same with the mean string length from the caller (lanes per pair follow it: 1 lane per 16 bytes, up to a warp;
 * ta_hamming_batch_dev cannot read the device offsets and assumes 64 bytes).  Any value gives the same results. */
int ta_hamming_batch_dev_len(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                             const uint64_t *b_off, size_t n, uint32_t mean_len, uint32_t *out, void *stream);
int ta_levenshtein_k_batch_dev(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                               const uint64_t *b_off, size_t n, uint32_t k, ta_costs costs, uint32_t max_len,
                               uint32_t *out, void *stream);
/* Exponential-k driver on device-resident pairs.  Synchronises `stream` between rounds (it has to read how many
 * pairs are still TA_NONE); results are complete in `out` when it returns. */
int ta_levenshtein_exp_batch_dev(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                                 const uint64_t *b_off, size_t n, ta_costs costs, uint32_t max_len, uint32_t *out,
                                 void *stream);
/* Search over device-resident haystacks (`needle` is a HOST pointer, needle_len >= 1; `max_hay_len` bounds the
 * haystack lengths).  Match arrays are returned in host memory exactly as by ta_levenshtein_search_batch. */
int ta_levenshtein_search_batch_dev(ta_ctx *ctx, const uint8_t *needle, size_t needle_len, const uint8_t *hay,
                                    const uint64_t *hay_off, size_t n, uint64_t max_hay_len, uint32_t k,
                                    int search_type, ta_costs costs, int anchored, ta_match **out_matches,
                                    uint64_t **out_match_off, void *stream);
/* Synchronises `stream` and reports a deferred contract violation seen by a *_dev kernel (e.g. a Hamming
 * length mismatch), clearing it. */
int ta_dev_status(ta_ctx *ctx, void *stream);

/* ---- single-pair conveniences with the crate's exact shapes (batch of one through the same kernels) -------- */
int ta_hamming(ta_ctx *ctx, const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len, uint32_t *out);
int ta_levenshtein_simd_k_with_opts(ta_ctx *ctx, const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len,
                                    uint32_t k, ta_costs costs, uint32_t *out);
int ta_levenshtein_exp_with_opts(ta_ctx *ctx, const uint8_t *a, size_t a_len, const uint8_t *b, size_t b_len,
                                 ta_costs costs, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* TRIPLE_ACCEL_B200_H */
