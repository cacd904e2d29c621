// lev_bitpar.cu -- bit-parallel fast paths for unit costs (LEVENSHTEIN_COSTS / RDAMERAU_COSTS) on sm_100a.
//
//  * lev_bitpar_tab_kernel<TRANS, PLANES, W> / lev_bitpar32_kernel: k-bounded distance, one thread per pair, band of <= 32
//    diagonals with the 32-bit window (k <= 31, or <= 30 with transpositions) and <= 64 with the 64-bit window.  Algorithm and data movement are described in
//    lev_bitpar_core.cuh (match-table variant = default; SWAR-compare variant = register-only alternative).  Same
//    contract as the general kernel in lev_band.cu (reference src/levenshtein.rs:376-545); the dispatcher in api.cu
//    picks these whenever the cost model is unit and the band fits.
//  * search_filter_kernel: (below) flags haystacks that contain at least one match end with cost <= k.
#include <stdlib.h>

#include <algorithm>

#include "lev_bitpar_core.cuh"
#include "ta_common.cuh"

// defaults of the block-table kernel (measured on B200, profiles/README.md)
#define TA_BLK_DEFAULT_PLANES 1

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

// Table variant (lev_bitpar_core.cuh: distance32_tab): persistent threads, each with a private 128-entry column of
// the block's shared-memory match table (layout [entry][thread]: bank = thread, conflict-free for any bytes).
template <bool TRANS, int PLANES, typename W>
__global__ void __launch_bounds__(128) lev_bitpar_tab_kernel(const uint8_t *__restrict__ a,
                                                               const uint64_t *__restrict__ a_off,
                                                               const uint8_t *__restrict__ b,
                                                               const uint64_t *__restrict__ b_off,
                                                               const uint32_t *__restrict__ idx, size_t n, uint32_t k,
                                                               uint32_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t tabs_raw[];
    W *tabs = (W *)tabs_raw;
    const uint32_t nt = blockDim.x;
    for (uint32_t q = threadIdx.x; q < (128u >> (PLANES - 1)) * nt; q += nt) tabs[q] = 0;
    __syncthreads();
    uint8_t *tab = (uint8_t *)(tabs + threadIdx.x);
    const uint32_t pitch = nt * (uint32_t)sizeof(W);  // bytes between consecutive entries of one thread
    const size_t total = (size_t)gridDim.x * nt;
    for (size_t w = (size_t)blockIdx.x * nt + threadIdx.x; w < n; w += total) {
        const size_t pair = idx ? (size_t)idx[w] : w;
        const uint64_t a0 = a_off[pair], a1 = a_off[pair + 1];
        const uint64_t b0 = b_off[pair], b1 = b_off[pair + 1];
        out[pair] = bitpar::pair_unit_costs_tab<TRANS, PLANES, W>(a + a0, a1 - a0, b + b0, b1 - b0, k, tab, pitch);
    }
}


// Block-table variant (lev_bitpar_core.cuh: distance_blk), bands of <= 33 - C diagonals.  Same table layout
// ([entry][thread] u32, 128 or 256 entries per thread).  The persistent pair loop is software-pipelined: the CSR
// offsets are loaded two pairs ahead and the first lines of the next pair's strings are pulled into L2 one pair
// ahead, so a pair never starts with two dependent DRAM round trips (offsets -> bytes).
struct PairRef {
    uint64_t a0, b0;
    uint32_t alen, blen;
    size_t pair;
};
__device__ __forceinline__ PairRef load_pair_ref(const uint64_t *__restrict__ a_off, const uint64_t *__restrict__ b_off,
                                                 const uint32_t *__restrict__ idx, size_t w, size_t n) {
    PairRef r;
    r.a0 = r.b0 = 0, r.alen = r.blen = 0, r.pair = 0;
    if (w < n) {
        r.pair = idx ? (size_t)idx[w] : w;
        r.a0 = a_off[r.pair];
        r.b0 = b_off[r.pair];
        r.alen = (uint32_t)(a_off[r.pair + 1] - r.a0);
        r.blen = (uint32_t)(b_off[r.pair + 1] - r.b0);
    }
    return r;
}
__device__ __forceinline__ void prefetch_l2(const uint8_t *p, uint32_t len) {
    if (len == 0) return;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    if (((uintptr_t)p & 127) + len > 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 128));
}

template <bool TRANS, int PLANES, int C>
__global__ void __launch_bounds__(256) lev_bitpar_blk_kernel(const uint8_t *__restrict__ a,
                                                             const uint64_t *__restrict__ a_off,
                                                             const uint8_t *__restrict__ b,
                                                             const uint64_t *__restrict__ b_off,
                                                             const uint32_t *__restrict__ idx, size_t n, uint32_t k,
                                                             uint32_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t tabs_raw[];
    uint32_t *tabs = (uint32_t *)tabs_raw;
    const uint32_t nt = blockDim.x;
    for (uint32_t q = threadIdx.x; q < (PLANES ? 128u : 256u) * nt; q += nt) tabs[q] = 0;
    __syncthreads();
    uint8_t *tab = (uint8_t *)(tabs + threadIdx.x);
    const uint32_t pitch = nt * 4u;
    const size_t total = (size_t)gridDim.x * nt;
    size_t w = (size_t)blockIdx.x * nt + threadIdx.x;
    PairRef cur = load_pair_ref(a_off, b_off, idx, w, n);
    PairRef nxt = load_pair_ref(a_off, b_off, idx, w + total, n);
    for (; w < n; w += total) {
        const PairRef nn = load_pair_ref(a_off, b_off, idx, w + 2 * total, n);  // in flight during this pair
        prefetch_l2(a + nxt.a0, nxt.alen);
        prefetch_l2(b + nxt.b0, nxt.blen);
        bitpar::NextHint hint;  // first bytes of the next pair: into L1 when this pair leaves its main loop
        hint.p[0] = nxt.alen ? a + nxt.a0 : nullptr;
        hint.p[1] = nxt.blen ? b + nxt.b0 : nullptr;
        hint.p[2] = hint.p[3] = nullptr;
        hint.len[0] = nxt.alen, hint.len[1] = nxt.blen, hint.len[2] = hint.len[3] = 0;
        out[cur.pair] = bitpar::pair_unit_costs_blk<TRANS, PLANES, C>(a + cur.a0, cur.alen, b + cur.b0, cur.blen, k, tab, pitch, &hint);
        cur = nxt;
        nxt = nn;
    }
}

// Two pairs per thread (lev_bitpar_core.cuh: distance_duo): bands of <= 9 diagonals, unit costs without
// transpositions.  Work item w = pairs 2w and 2w + 1; when the two cannot share a recurrence (different numbers of
// 16-column supersteps, an answer known without DP, an odd last pair) each goes through the single-pair routine.
__global__ void __launch_bounds__(128, 3) lev_bitpar_duo_kernel(const uint8_t *__restrict__ a,
                                                                const uint64_t *__restrict__ a_off,
                                                                const uint8_t *__restrict__ b,
                                                                const uint64_t *__restrict__ b_off,
                                                                const uint32_t *__restrict__ idx, size_t n, uint32_t k,
                                                                uint32_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t tabs_raw[];
    uint32_t *tabs = (uint32_t *)tabs_raw;
    const uint32_t nt = blockDim.x;
    for (uint32_t q = threadIdx.x; q < 128u * nt; q += nt) tabs[q] = 0;
    __syncthreads();
    uint8_t *tab = (uint8_t *)(tabs + threadIdx.x);
    const uint32_t pitch = nt * 4u;
    const size_t total = (size_t)gridDim.x * nt;
    const size_t items = (n + 1) / 2;
    size_t w = (size_t)blockIdx.x * nt + threadIdx.x;
    PairRef cur0 = load_pair_ref(a_off, b_off, idx, 2 * w, n), cur1 = load_pair_ref(a_off, b_off, idx, 2 * w + 1, n);
    for (; w < items; w += total) {
        const size_t wn = w + total;
        const PairRef nx0 = load_pair_ref(a_off, b_off, idx, 2 * wn, n);  // in flight during this item
        const PairRef nx1 = load_pair_ref(a_off, b_off, idx, 2 * wn + 1, n);
        const bool has1 = 2 * w + 1 < n;
        const uint8_t *pa0 = a + cur0.a0, *pb0 = b + cur0.b0, *pa1 = a + cur1.a0, *pb1 = b + cur1.b0;
        uint64_t la0 = cur0.alen, lb0 = cur0.blen, la1 = cur1.alen, lb1 = cur1.blen;
        uint32_t mk0 = 0, mk1 = 0, r0 = 0, r1 = 0;
        const bool dp0 = bitpar::unit_costs_prepare(pa0, la0, pb0, lb0, k, mk0, &r0);
        const bool dp1 = has1 && bitpar::unit_costs_prepare(pa1, la1, pb1, lb1, k, mk1, &r1);
        // The choice is made per WARP: if any lane's two pairs cannot share a recurrence, every lane of the warp takes
        // the single-pair path.  A warp that runs both paths pays for both one after the other (and drags both code
        // regions through the instruction cache): on ragged batches that was 2.4x slower than never pairing at all.
        const bool can_pair = dp0 && dp1 && (lb0 >> 4) == (lb1 >> 4);
        const bool skip = !dp0 && !dp1;  // nothing to compute for this lane
        if (__all_sync(__activemask(), can_pair || skip) && can_pair) {
            bitpar::NextHint hint;  // first bytes of the next item's four strings: into L1 at the end of the main loop
            hint.p[0] = nx0.alen ? a + nx0.a0 : nullptr;
            hint.p[1] = nx0.blen ? b + nx0.b0 : nullptr;
            hint.p[2] = nx1.alen ? a + nx1.a0 : nullptr;
            hint.p[3] = nx1.blen ? b + nx1.b0 : nullptr;
            hint.len[0] = nx0.alen, hint.len[1] = nx0.blen, hint.len[2] = nx1.alen, hint.len[3] = nx1.blen;
            bitpar::distance_duo(pa0, (int)la0, pb0, (int)lb0, mk0, pa1, (int)la1, pb1, (int)lb1, mk1, tab, pitch, r0, r1,
                                 &hint);
            r0 = r0 <= mk0 ? r0 : 0xFFFFFFFFu;
            r1 = r1 <= mk1 ? r1 : 0xFFFFFFFFu;
        } else {
            // one pair after the other through ONE inlined copy of the single-pair routine: the kernel's hot code has
            // to stay within the 32 KB L1.5 instruction cache (two copies made every warp on an SM thrash it)
#pragma unroll 1
            for (int q = 0; q < 2; q++) {
                if (q == 0 ? dp0 : dp1) {
                    const uint32_t mk = q == 0 ? mk0 : mk1;
                    uint32_t r = bitpar::distance_blk<false, 1, 16>(q == 0 ? pa0 : pa1, (int)(q == 0 ? la0 : la1),
                                                                    q == 0 ? pb0 : pb1, (int)(q == 0 ? lb0 : lb1), mk, tab, pitch);
                    r = r <= mk ? r : 0xFFFFFFFFu;
                    if (q == 0)
                        r0 = r;
                    else
                        r1 = r;
                }
            }
        }
        out[cur0.pair] = r0;
        if (has1) out[cur1.pair] = r1;
        cur0 = nx0;
        cur1 = nx1;
    }
}


// The same two-pairs-per-thread routine behind a tile-local ordering (round 2; taken for ragged batches, see
// ta_set_length_hint; TA_DUO_TILED=0|1 forces one of the two kernels).  A thread's two pairs can only share a recurrence when they have the same number of 16-column supersteps, and
// a warp runs as long as its longest pair -- on ragged batches (lengths 96..160) every warp fell back to the single-pair
// path and waited for its longest lane: 0.32 ms per 1 M pairs against 0.20 ms for equal lengths.  A global counting
// sort fixes that (TA_LEN_BUCKETS=1, below) but costs three launches, ~30 us, on every batch.  Here each CTA takes a
// TILE of DUO_TILE consecutive pairs, counts them by length class (longer length / 16) in shared memory, and -- unless
// one class holds the whole tile, the case of equal-length batches, which keep their order -- places the tile's pair
// indices in class order (2 bytes per pair).  Round r of the tile then gives thread t the sorted positions
// 2 (128 r + t) and + 1: a thread's two pairs and a warp's 64 pairs are neighbours in length.  The pass costs ~12
// instructions per pair (a pair is ~4 000) and the offsets it reads are read again, from L1 / L2, when the pair is run.
constexpr int DUO_TILE = 3072;   // most pairs per tile (2-byte slots in shared memory); a multiple of 256 = one round
constexpr int DUO_CLASSES = 64;  // length classes (longer length / 16, clamped)
constexpr int DUO_SLOTS = DUO_TILE + DUO_CLASSES;  // slots of a tile's order: every class starts at an even position
__global__ void __launch_bounds__(128, 3) lev_bitpar_duo_tiled_kernel(const uint8_t *__restrict__ a,
                                                                      const uint64_t *__restrict__ a_off,
                                                                      const uint8_t *__restrict__ b,
                                                                      const uint64_t *__restrict__ b_off,
                                                                      const uint32_t *__restrict__ idx, size_t n, uint32_t k,
                                                                      uint16_t *__restrict__ order_ws,
                                                                      uint32_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t tabs_raw[];
    uint32_t *tabs = (uint32_t *)tabs_raw;
    const uint32_t nt = blockDim.x;  // 128
    // Not one byte of shared memory beyond the match tables: 3 CTAs x 64 KB fit the 196 KB carve-out, 3 x 70 KB need the
    // 228 KB one, and the 32 KB of L1 that costs made every batch 15 % slower (the streams of 12 warps x 32 lanes live
    // in L1 between their 16-byte reads).  So the tile's order goes to a global scratch slice of this CTA (2 bytes per
    // pair, read back through L1) and the class counters borrow the first 64 words of the tables, which are all-zero
    // between pairs and are zeroed again before the first pair of the tile.
    uint16_t *order = order_ws + (size_t)blockIdx.x * DUO_SLOTS;  // tile-local pair indices in class order
    uint32_t *hist = tabs;                                        // [DUO_CLASSES] counts, then write cursors
    volatile uint32_t *tile_ident = tabs + DUO_CLASSES;           // two more borrowed words: identity flag, slots in use
    uint8_t *tab = (uint8_t *)(tabs + threadIdx.x);
    const uint32_t pitch = nt * 4u;
    bool tables_clean = false;
    // Work split: the batch is cut into rounds of 256 pairs (128 threads x 2); CTA b takes a contiguous range of rounds,
    // every CTA within one round of the others (an equal share rounded UP to whole rounds left 3 of 148 SMs idle), and cuts
    // its range into tiles of <= DUO_TILE pairs.
    const size_t rounds_total = (n + 255) / 256;
    const size_t rq = rounds_total / gridDim.x, rrem = rounds_total % gridDim.x;
    const size_t r_lo = (size_t)blockIdx.x * rq + (blockIdx.x < rrem ? blockIdx.x : rrem);
    const size_t r_cnt = rq + (blockIdx.x < rrem ? 1 : 0);
    const size_t n_tiles = (r_cnt + DUO_TILE / 256 - 1) / (DUO_TILE / 256);
    auto len_class_of = [&](size_t w) -> uint32_t {
        const size_t pair = idx ? (size_t)idx[w] : w;
        const uint64_t la = a_off[pair + 1] - a_off[pair], lb = b_off[pair + 1] - b_off[pair];
        const uint64_t c = (la > lb ? la : lb) >> 4;
        return c < (uint64_t)(DUO_CLASSES - 1) ? (uint32_t)c : (uint32_t)(DUO_CLASSES - 1);
    };
    for (size_t tile = 0; tile < n_tiles; tile++) {
        const size_t t_lo = r_lo + r_cnt * tile / n_tiles, t_hi = r_lo + r_cnt * (tile + 1) / n_tiles;  // rounds
        const size_t base = t_lo * 256;
        const size_t end = t_hi * 256 < n ? t_hi * 256 : n;
        const uint32_t cnt = (uint32_t)(end - base);
        // ---- order the tile by length class ----------------------------------------------------------------------
        // classes of pairs tid + 128 j (j < 24) as bytes of pk[] (0xFF = none); the offsets are fetched 12 pairs at a time
        // so that a thread waits for two round trips per tile, not for one per pair (that was 18 us per pass), and the
        // first fetch of the kernel is in flight while the match tables are zeroed.  (Measured against this: thread t
        // taking 18 CONSECUTIVE pairs -- 50 loads of neighbouring low words, classes from register differences, ~150
        // instructions instead of ~900 -- was 7 % slower on every workload; the whole kernel is sensitive to where its
        // unrolled body lands in the instruction cache.)
        uint32_t pk[DUO_TILE / 128 / 4];
#pragma unroll
        for (int q = 0; q < DUO_TILE / 128 / 4; q++) pk[q] = 0xFFFFFFFFu;
#pragma unroll
        for (int g = 0; g < 2; g++) {
            uint32_t c[12];
#pragma unroll
            for (int u = 0; u < 12; u++) {
                const uint32_t i = threadIdx.x + 128u * (uint32_t)(g * 12 + u);
                c[u] = i < cnt ? len_class_of(base + i) : 0xFFu;
            }
            if (g == 0) {
                if (!tables_clean) {
                    for (uint32_t q = threadIdx.x; q < 128u * nt; q += nt) tabs[q] = 0;
                    tables_clean = true;
                }
                __syncthreads();  // tables (and with them the borrowed counters) are zero
            }
#pragma unroll
            for (int u = 0; u < 12; u++) {
                const int j = g * 12 + u;
                // equal-length batches: the 32 lanes of a warp count into the same word -- one atomic for the warp
                const uint32_t c_first = __shfl_sync(0xffffffffu, c[u], 0);
                if (__all_sync(0xffffffffu, c[u] == c_first)) {
                    if ((threadIdx.x & 31u) == 0 && c_first != 0xFFu) atomicAdd(&hist[c_first], 32u);
                } else if (c[u] != 0xFFu) {
                    atomicAdd(&hist[c[u]], 1u);
                }
                pk[j >> 2] = (pk[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | (c[u] << (8 * (j & 3)));
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {  // one warp: is the tile one class?  else exclusive scan of the 64 counts -> cursors
            const uint32_t c0 = hist[threadIdx.x], c1 = hist[threadIdx.x + 32];
            const bool one = __any_sync(0xffffffffu, c0 == cnt || c1 == cnt);
            // every class starts at an EVEN position (a thread takes positions 2s and 2s + 1, so its two pairs are always
            // of one class); a class with an odd count leaves its last slot empty (0xFFFF) and that thread lends its one
            // pair to the idle half.  A thread straddling two classes sent its whole warp down the single-pair path,
            // and that code evicted the paired path from the instruction cache of its SM.
            const uint32_t e0 = (c0 + 1u) & ~1u, e1 = (c1 + 1u) & ~1u;
            uint32_t s0 = e0, s1 = e1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t0 = __shfl_up_sync(0xffffffffu, s0, o), t1 = __shfl_up_sync(0xffffffffu, s1, o);
                if ((int)threadIdx.x >= o) s0 += t0, s1 += t1;
            }
            const uint32_t tot0 = __shfl_sync(0xffffffffu, s0, 31), tot1 = __shfl_sync(0xffffffffu, s1, 31);
            const uint32_t b0 = s0 - e0, b1 = tot0 + s1 - e1;  // first slot of classes t and t + 32
            hist[threadIdx.x] = b0;
            hist[threadIdx.x + 32] = b1;
            if (!one) {
                if (c0 & 1u) order[b0 + c0] = 0xFFFFu;
                if (c1 & 1u) order[b1 + c1] = 0xFFFFu;
            }
            if (threadIdx.x == 0) {
                *tile_ident = one ? 1u : 0u;
                tile_ident[1] = one ? cnt : tot0 + tot1;  // slots in use
            }
        }
        __syncthreads();
        const bool ident = *tile_ident != 0u;
        const uint32_t plen = tile_ident[1];
        if (!ident) {
#pragma unroll
            for (int j = 0; j < DUO_TILE / 128; j++) {
                const uint32_t cj = (pk[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                if (cj != 0xFFu) order[atomicAdd(&hist[cj], 1u)] = (uint16_t)(threadIdx.x + 128u * (uint32_t)j);
            }
            __threadfence_block();  // the slots are read by other threads of the CTA
            __syncthreads();
        }
        __syncthreads();                                           // everyone has read tile_ident
        if (threadIdx.x <= DUO_CLASSES + 1) hist[threadIdx.x] = 0;  // hand the borrowed words back to the tables
        __syncthreads();
        // ---- the rounds of the tile ------------------------------------------------------------------------------
        // sorted position -> pair (valid = false: an empty slot or past the end)
        auto ref_at = [&](uint32_t pos, bool &valid) -> PairRef {
            uint32_t local = pos;
            valid = pos < plen;
            if (valid && !ident) {
                local = (uint32_t)((volatile uint16_t *)order)[pos];
                valid = local != 0xFFFFu;
            }
            return load_pair_ref(a_off, b_off, idx, valid ? base + local : n, n);
        };
        uint32_t slot = threadIdx.x;
        bool has0, has1, nh0, nh1;
        PairRef cur0 = ref_at(2 * slot, has0), cur1 = ref_at(2 * slot + 1, has1);
        for (; 2 * (slot - threadIdx.x) < plen; slot += nt) {
            const PairRef nx0 = ref_at(2 * (slot + nt), nh0), nx1 = ref_at(2 * (slot + nt) + 1, nh1);  // in flight during this item
            const uint8_t *pa0 = a + cur0.a0, *pb0 = b + cur0.b0, *pa1 = a + cur1.a0, *pb1 = b + cur1.b0;
            uint64_t la0 = cur0.alen, lb0 = cur0.blen, la1 = cur1.alen, lb1 = cur1.blen;
            uint32_t mk0 = 0, mk1 = 0, r0 = 0, r1 = 0;
            const bool dp0 = has0 && bitpar::unit_costs_prepare(pa0, la0, pb0, lb0, k, mk0, &r0);
            const bool dp1 = has1 && bitpar::unit_costs_prepare(pa1, la1, pb1, lb1, k, mk1, &r1);
            // A lane with only ONE pair to compute (odd class count at a class boundary, the last pair of the batch, a
            // partner whose answer needs no DP) lends it to the other half and discards that result: the warp stays on
            // the paired path instead of taking the single-pair path for all its lanes.
            if (dp0 && !dp1) pa1 = pa0, pb1 = pb0, la1 = la0, lb1 = lb0, mk1 = mk0;
            if (dp1 && !dp0) pa0 = pa1, pb0 = pb1, la0 = la1, lb0 = lb1, mk0 = mk1;
            // the choice is made per WARP, as in the kernel above
            const bool can_pair = (dp0 || dp1) && (lb0 >> 4) == (lb1 >> 4);
            const bool skip = !dp0 && !dp1;
            if (__all_sync(__activemask(), can_pair || skip) && can_pair) {
                bitpar::NextHint hint;
                hint.p[0] = nx0.alen ? a + nx0.a0 : nullptr;
                hint.p[1] = nx0.blen ? b + nx0.b0 : nullptr;
                hint.p[2] = nx1.alen ? a + nx1.a0 : nullptr;
                hint.p[3] = nx1.blen ? b + nx1.b0 : nullptr;
                hint.len[0] = nx0.alen, hint.len[1] = nx0.blen, hint.len[2] = nx1.alen, hint.len[3] = nx1.blen;
                uint32_t t0 = 0, t1 = 0;
                bitpar::distance_duo(pa0, (int)la0, pb0, (int)lb0, mk0, pa1, (int)la1, pb1, (int)lb1, mk1, tab, pitch, t0, t1,
                                     &hint);
                if (dp0) r0 = t0 <= mk0 ? t0 : 0xFFFFFFFFu;
                if (dp1) r1 = t1 <= mk1 ? t1 : 0xFFFFFFFFu;
            } else {
#pragma unroll 1
                for (int q = 0; q < 2; q++) {
                    if (q == 0 ? dp0 : dp1) {
                        const uint32_t mk = q == 0 ? mk0 : mk1;
                        uint32_t rr = bitpar::distance_blk<false, 1, 16>(q == 0 ? pa0 : pa1, (int)(q == 0 ? la0 : la1),
                                                                         q == 0 ? pb0 : pb1, (int)(q == 0 ? lb0 : lb1), mk, tab, pitch);
                        rr = rr <= mk ? rr : 0xFFFFFFFFu;
                        if (q == 0)
                            r0 = rr;
                        else
                            r1 = rr;
                    }
                }
            }
            if (has0) out[cur0.pair] = r0;
            if (has1) out[cur1.pair] = r1;
            cur0 = nx0, cur1 = nx1;
            has0 = nh0, has1 = nh1;
        }
        __syncthreads();  // `order` and the borrowed words are rewritten for the next tile
    }
}

// The block-table kernel (one pair per thread) behind the same tile-local ordering, for ragged batches (bands of 10 ..
// 25 diagonals, transpositions): a warp runs as long as its longest pair, and in class order its 32 pairs are neighbours
// in length.  Same bookkeeping as lev_bitpar_duo_tiled_kernel: contiguous rounds (of 128 pairs here) per CTA, tiles of
// <= DUO_TILE pairs, class counters borrowed from the clean match tables, the order in a global scratch slice.
template <bool TRANS, int C>
__global__ void __launch_bounds__(128) lev_bitpar_blk_tiled_kernel(const uint8_t *__restrict__ a,
                                                                   const uint64_t *__restrict__ a_off,
                                                                   const uint8_t *__restrict__ b,
                                                                   const uint64_t *__restrict__ b_off,
                                                                   const uint32_t *__restrict__ idx, size_t n, uint32_t k,
                                                                   uint16_t *__restrict__ order_ws,
                                                                   uint32_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t tabs_raw[];
    uint32_t *tabs = (uint32_t *)tabs_raw;
    const uint32_t nt = blockDim.x;  // 128
    uint16_t *order = order_ws + (size_t)blockIdx.x * DUO_SLOTS;
    uint32_t *hist = tabs;
    volatile uint32_t *tile_ident = tabs + DUO_CLASSES;
    uint8_t *tab = (uint8_t *)(tabs + threadIdx.x);
    const uint32_t pitch = nt * 4u;
    bool tables_clean = false;
    const size_t rounds_total = (n + 127) / 128;
    const size_t rq = rounds_total / gridDim.x, rrem = rounds_total % gridDim.x;
    const size_t r_lo = (size_t)blockIdx.x * rq + (blockIdx.x < rrem ? blockIdx.x : rrem);
    const size_t r_cnt = rq + (blockIdx.x < rrem ? 1 : 0);
    const size_t n_tiles = (r_cnt + DUO_TILE / 128 - 1) / (DUO_TILE / 128);
    auto len_class_of = [&](size_t w) -> uint32_t {
        const size_t pair = idx ? (size_t)idx[w] : w;
        const uint64_t la = a_off[pair + 1] - a_off[pair], lb = b_off[pair + 1] - b_off[pair];
        const uint64_t c = (la > lb ? la : lb) >> 4;
        return c < (uint64_t)(DUO_CLASSES - 1) ? (uint32_t)c : (uint32_t)(DUO_CLASSES - 1);
    };
    for (size_t tile = 0; tile < n_tiles; tile++) {
        const size_t t_lo = r_lo + r_cnt * tile / n_tiles, t_hi = r_lo + r_cnt * (tile + 1) / n_tiles;  // rounds
        const size_t base = t_lo * 128;
        const size_t end = t_hi * 128 < n ? t_hi * 128 : n;
        const uint32_t cnt = (uint32_t)(end - base);
        uint32_t pk[DUO_TILE / 128 / 4];
#pragma unroll
        for (int q = 0; q < DUO_TILE / 128 / 4; q++) pk[q] = 0xFFFFFFFFu;
#pragma unroll
        for (int g = 0; g < 2; g++) {
            uint32_t c[12];
#pragma unroll
            for (int u = 0; u < 12; u++) {
                const uint32_t i = threadIdx.x + 128u * (uint32_t)(g * 12 + u);
                c[u] = i < cnt ? len_class_of(base + i) : 0xFFu;
            }
            if (g == 0) {
                if (!tables_clean) {
                    for (uint32_t q = threadIdx.x; q < 128u * nt; q += nt) tabs[q] = 0;
                    tables_clean = true;
                }
                __syncthreads();
            }
#pragma unroll
            for (int u = 0; u < 12; u++) {
                const int j = g * 12 + u;
                const uint32_t c_first = __shfl_sync(0xffffffffu, c[u], 0);
                if (__all_sync(0xffffffffu, c[u] == c_first)) {
                    if ((threadIdx.x & 31u) == 0 && c_first != 0xFFu) atomicAdd(&hist[c_first], 32u);
                } else if (c[u] != 0xFFu) {
                    atomicAdd(&hist[c[u]], 1u);
                }
                pk[j >> 2] = (pk[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | (c[u] << (8 * (j & 3)));
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t c0 = hist[threadIdx.x], c1 = hist[threadIdx.x + 32];
            const bool one = __any_sync(0xffffffffu, c0 == cnt || c1 == cnt);
            uint32_t s0 = c0, s1 = c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t0 = __shfl_up_sync(0xffffffffu, s0, o), t1 = __shfl_up_sync(0xffffffffu, s1, o);
                if ((int)threadIdx.x >= o) s0 += t0, s1 += t1;
            }
            const uint32_t tot0 = __shfl_sync(0xffffffffu, s0, 31);
            hist[threadIdx.x] = s0 - c0;
            hist[threadIdx.x + 32] = tot0 + s1 - c1;
            if (threadIdx.x == 0) *tile_ident = one ? 1u : 0u;
        }
        __syncthreads();
        const bool ident = *tile_ident != 0u;
        if (!ident) {
#pragma unroll
            for (int j = 0; j < DUO_TILE / 128; j++) {
                const uint32_t cj = (pk[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                if (cj != 0xFFu) order[atomicAdd(&hist[cj], 1u)] = (uint16_t)(threadIdx.x + 128u * (uint32_t)j);
            }
        }
        __syncthreads();
        if (threadIdx.x <= DUO_CLASSES) hist[threadIdx.x] = 0;  // hand the borrowed words back to the tables
        __syncthreads();
        auto ref_at = [&](uint32_t pos) -> PairRef {  // sorted position -> pair (w = n: none)
            if (pos >= cnt) return load_pair_ref(a_off, b_off, idx, n, n);
            return load_pair_ref(a_off, b_off, idx, base + (ident ? pos : (uint32_t)((volatile uint16_t *)order)[pos]), n);
        };
        uint32_t slot = threadIdx.x;
        PairRef cur = ref_at(slot), nxt = ref_at(slot + nt);
        for (; slot - threadIdx.x < cnt; slot += nt) {
            const PairRef nn = ref_at(slot + 2 * nt);  // in flight during this pair
            prefetch_l2(a + nxt.a0, nxt.alen);
            prefetch_l2(b + nxt.b0, nxt.blen);
            bitpar::NextHint hint;
            hint.p[0] = nxt.alen ? a + nxt.a0 : nullptr;
            hint.p[1] = nxt.blen ? b + nxt.b0 : nullptr;
            hint.p[2] = hint.p[3] = nullptr;
            hint.len[0] = nxt.alen, hint.len[1] = nxt.blen, hint.len[2] = hint.len[3] = 0;
            if (slot < cnt)
                out[cur.pair] = bitpar::pair_unit_costs_blk<TRANS, 1, C>(a + cur.a0, cur.alen, b + cur.b0, cur.blen, k, tab, pitch, &hint);
            cur = nxt;
            nxt = nn;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Length bucketing (experimental, TA_LEN_BUCKETS=1): a counting sort of the pair indices by (longer length) / 16, so
// that neighbouring work items have the same number of 16-column supersteps -- pairs can then share a thread in the
// duo kernel and the lanes of a warp finish together -- handed to the distance kernels as the index indirection they
// already take for the exponential rounds.  Three small launches; off by default until measured.
__device__ __forceinline__ uint32_t len_class(const uint64_t *__restrict__ a_off, const uint64_t *__restrict__ b_off, size_t i) {
    const uint64_t la = a_off[i + 1] - a_off[i], lb = b_off[i + 1] - b_off[i];
    const uint64_t c = (la > lb ? la : lb) >> 4;
    return c < 255 ? (uint32_t)c : 255u;
}
__global__ void __launch_bounds__(256) len_hist_kernel(const uint64_t *__restrict__ a_off, const uint64_t *__restrict__ b_off,
                                                       size_t n, uint32_t *__restrict__ hist) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        atomicAdd(&sh[len_class(a_off, b_off, i)], 1u);
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void len_scan_kernel(const uint32_t *__restrict__ hist, uint32_t *__restrict__ pos) {
    if (threadIdx.x == 0) {
        uint32_t sum = 0;
        for (int c = 0; c < 256; c++) {
            pos[c] = sum;
            sum += hist[c];
        }
    }
}
__global__ void __launch_bounds__(256) len_scatter_kernel(const uint64_t *__restrict__ a_off,
                                                          const uint64_t *__restrict__ b_off, size_t n,
                                                          uint32_t *__restrict__ pos, uint32_t *__restrict__ idx) {
    __shared__ uint32_t cnt[256], base[256];
    // each pass of the block places blockDim.x pairs: count per class, reserve a global range per class, then place
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x; i0 < n; i0 += (size_t)gridDim.x * blockDim.x) {
        cnt[threadIdx.x] = 0;
        __syncthreads();
        const size_t i = i0 + threadIdx.x;
        uint32_t c = 0, rank = 0;
        if (i < n) {
            c = len_class(a_off, b_off, i);
            rank = atomicAdd(&cnt[c], 1u);
        }
        __syncthreads();
        if (cnt[threadIdx.x]) base[threadIdx.x] = atomicAdd(&pos[threadIdx.x], cnt[threadIdx.x]);
        __syncthreads();
        if (i < n) idx[base[c] + rank] = (uint32_t)i;
        __syncthreads();
    }
}

typedef void (*lev_kern_t)(const uint8_t *, const uint64_t *, const uint8_t *, const uint64_t *, const uint32_t *,
                           size_t, uint32_t, uint32_t *);
template <bool TRANS, int C>
lev_kern_t pick_blk(int planes) {
    return planes ? lev_bitpar_blk_kernel<TRANS, 1, C> : lev_bitpar_blk_kernel<TRANS, 0, C>;
}

}  // namespace

bool ta_bitpar_can_handle(uint32_t k, ta_costs c, uint32_t max_len) {
    if (!(c.mismatch == 1 && c.gap == 1 && c.start_gap == 0 && c.transpose <= 1)) return false;
    const uint32_t kk = k < max_len ? k : max_len;  // max_k = min(k, n) for unit costs
    return kk <= (c.transpose ? 62u : 63u);  // band (+ transposition margin) <= 64 rows, see lev_bitpar_core.cuh
}
static bool fits32(uint32_t k, ta_costs c, uint32_t max_len) {
    const uint32_t kk = k < max_len ? k : max_len;
    return kk <= (c.transpose ? 30u : 31u);
}
static bool fits16(uint32_t k, ta_costs c, uint32_t max_len) {
    const uint32_t kk = k < max_len ? k : max_len;
    return kk <= (c.transpose ? 14u : 15u);
}

int ta_launch_lev_bitpar(ta_ctx *ctx, const uint8_t *a, const uint64_t *a_off, const uint8_t *b,
                         const uint64_t *b_off, size_t n, const uint32_t *idx, uint32_t k, ta_costs costs,
                         uint32_t max_len, uint32_t *out, cudaStream_t st) {
    if (n == 0) return TA_OK;
    static const int len_buckets = getenv("TA_LEN_BUCKETS") ? atoi(getenv("TA_LEN_BUCKETS")) : 0;
    if (len_buckets && idx == nullptr && n >= 4096 && n < 0xFFFFFFF0ull) {
        int rc;
        if ((rc = ta_dev_reserve(ctx, ctx->d_work[2], n * sizeof(uint32_t))) != TA_OK) return rc;
        if ((rc = ta_dev_reserve(ctx, ctx->d_work[3], 512 * sizeof(uint32_t))) != TA_OK) return rc;
        uint32_t *hist = (uint32_t *)ctx->d_work[3].p, *pos = hist + 256, *perm = (uint32_t *)ctx->d_work[2].p;
        TA_CUDA(ctx, cudaMemsetAsync(hist, 0, 256 * sizeof(uint32_t), st));
        const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
        len_hist_kernel<<<blocks, 256, 0, st>>>(a_off, b_off, n, hist);
        len_scan_kernel<<<1, 32, 0, st>>>(hist, pos);
        len_scatter_kernel<<<blocks, 256, 0, st>>>(a_off, b_off, n, pos, perm);
        ctx->launches += 3;
        TA_CUDA(ctx, cudaGetLastError());
        idx = perm;
    }
    // Implementations of the same per-pair contract: the match-table kernel (default; 32-row window, or 64-row window
    // for 32 <= k <= 63) and the register-only SWAR-compare kernel (32 rows).  TA_BITPAR=simd|tab, TA_BITPAR_PLANES and
    // TA_BITPAR_THREADS force variants (the tests pin every one of them to the oracle).
    static const char *variant = getenv("TA_BITPAR");
    static const int env_threads = getenv("TA_BITPAR_THREADS") ? atoi(getenv("TA_BITPAR_THREADS")) : 0;
    static const int env_planes = getenv("TA_BITPAR_PLANES") ? atoi(getenv("TA_BITPAR_PLANES")) : 0;
    static const int env_bits = getenv("TA_BITPAR_BITS") ? atoi(getenv("TA_BITPAR_BITS")) : 0;
    // block-table kernel (default whenever the band fits 33 - C diagonals): TA_BLK_PLANES=0|1 (256 entries / 128
    // entries + top-bit plane), TA_BLK_C=8 forces 8-position blocks on narrow bands
    static const int blk_planes = getenv("TA_BLK_PLANES") ? atoi(getenv("TA_BLK_PLANES")) : TA_BLK_DEFAULT_PLANES;
    static const int blk_c = getenv("TA_BLK_C") ? atoi(getenv("TA_BLK_C")) : 0;
    {
        const uint32_t kk = k < max_len ? k : max_len;
        const uint32_t wmax = kk + 1 + (costs.transpose ? 1u : 0u);  // widest band any pair of the batch can have
        const bool use_blk = !(variant && (variant[0] == 's' || variant[0] == 't')) && wmax <= 25;
        // bands of <= 9 diagonals without transpositions: two pairs per thread (TA_BLK_DUO=0 disables)
        static const int blk_duo = getenv("TA_BLK_DUO") ? atoi(getenv("TA_BLK_DUO")) : 1;
        if (use_blk && blk_duo && !costs.transpose && wmax <= 9 && blk_planes == 1 && blk_c == 0) {
            const int nt = 128;
            // tile-local ordering by length class when the batch is ragged (decided per batch by the entry points, see
            // ta_set_length_hint; TA_DUO_TILED=0|1 forces one kernel)
            static const int env_tiled = getenv("TA_DUO_TILED") ? atoi(getenv("TA_DUO_TILED")) : -1;
            if (env_tiled >= 0 ? env_tiled != 0 : ctx->batch_ragged) {
                const size_t smem = (size_t)128 * nt * 4;
                const size_t rounds = (n + 255) / 256;  // the kernel splits them evenly over its CTAs
                const unsigned blocks = (unsigned)std::min<size_t>(rounds, (size_t)ctx->sm_count * 3);
                int rc = ta_dev_reserve(ctx, ctx->d_work[4], (size_t)blocks * DUO_SLOTS * sizeof(uint16_t));
                if (rc != TA_OK) return rc;
                TA_CUDA(ctx, cudaFuncSetAttribute(lev_bitpar_duo_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                lev_bitpar_duo_tiled_kernel<<<blocks, nt, smem, st>>>(a, a_off, b, b_off, idx, n, k, (uint16_t *)ctx->d_work[4].p, out);
                ctx->launches++;
                TA_CUDA(ctx, cudaGetLastError());
                return TA_OK;
            }
            const size_t smem = (size_t)128 * nt * 4;
            const size_t items = (n + 1) / 2;
            const unsigned blocks = (unsigned)std::min<size_t>((items + nt - 1) / nt, (size_t)ctx->sm_count * 3);
            TA_CUDA(ctx, cudaFuncSetAttribute(lev_bitpar_duo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            lev_bitpar_duo_kernel<<<blocks, nt, smem, st>>>(a, a_off, b, b_off, idx, n, k, out);
            ctx->launches++;
            TA_CUDA(ctx, cudaGetLastError());
            return TA_OK;
        }
        if (use_blk) {
            const int C = (wmax <= 17 && blk_c != 8) ? 16 : 8;
            static const int env_blk_tiled = getenv("TA_BLK_TILED") ? atoi(getenv("TA_BLK_TILED")) : -1;
            if ((env_blk_tiled >= 0 ? env_blk_tiled != 0 : ctx->batch_ragged) && blk_planes == 1 && env_threads == 0) {
                // ragged batch: tile-local ordering by length class (see lev_bitpar_blk_tiled_kernel; TA_BLK_TILED=0|1 forces)
                const int tnt = 128;
                const size_t tsmem = (size_t)128 * tnt * 4;
                const size_t rounds = (n + 127) / 128;
                const unsigned tblocks = (unsigned)std::min<size_t>(rounds, (size_t)ctx->sm_count * 3);
                int rc = ta_dev_reserve(ctx, ctx->d_work[4], (size_t)tblocks * DUO_SLOTS * sizeof(uint16_t));
                if (rc != TA_OK) return rc;
                void (*tk)(const uint8_t *, const uint64_t *, const uint8_t *, const uint64_t *, const uint32_t *, size_t, uint32_t,
                           uint16_t *, uint32_t *);
                if (C == 16)
                    tk = costs.transpose ? lev_bitpar_blk_tiled_kernel<true, 16> : lev_bitpar_blk_tiled_kernel<false, 16>;
                else
                    tk = costs.transpose ? lev_bitpar_blk_tiled_kernel<true, 8> : lev_bitpar_blk_tiled_kernel<false, 8>;
                TA_CUDA(ctx, cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
                tk<<<tblocks, tnt, tsmem, st>>>(a, a_off, b, b_off, idx, n, k, (uint16_t *)ctx->d_work[4].p, out);
                ctx->launches++;
                TA_CUDA(ctx, cudaGetLastError());
                return TA_OK;
            }
            const int nt = env_threads ? env_threads : (blk_planes ? 128 : 224);
            const size_t smem = (size_t)(blk_planes ? 128 : 256) * nt * 4;
            const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)(ctx->smem_optin + 1024) / (smem + 1024)));
            const unsigned blocks = (unsigned)std::min<size_t>((n + nt - 1) / nt, (size_t)ctx->sm_count * per_sm);
            lev_kern_t kern;
            if (C == 16)
                kern = costs.transpose ? pick_blk<true, 16>(blk_planes) : pick_blk<false, 16>(blk_planes);
            else
                kern = costs.transpose ? pick_blk<true, 8>(blk_planes) : pick_blk<false, 8>(blk_planes);
            TA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<blocks, nt, smem, st>>>(a, a_off, b, b_off, idx, n, k, out);
            ctx->launches++;
            TA_CUDA(ctx, cudaGetLastError());
            return TA_OK;
        }
    }
    const bool wide = !fits32(k, costs, max_len);
    if (wide || !(variant && variant[0] == 's')) {
        // window width: the narrowest that holds the band (16-bit entries halve the shared memory per thread and
        // double the resident warps); TA_BITPAR_BITS=32 keeps the 32-row window for narrow bands (measurement)
        const int bits = wide ? 64 : (fits16(k, costs, max_len) && env_bits != 32 ? 16 : 32);
        const int planes = bits == 64 ? 2 : (env_planes == 2 && bits == 32 ? 2 : 1);
        const int nt = env_threads ? env_threads : (bits == 64 ? 64 : 128);
        const size_t smem = (size_t)(planes == 2 ? 64 : 128) * nt * (bits / 8);
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)(ctx->smem_optin + 1024) / (smem + 1024)));
        const unsigned blocks = (unsigned)std::min<size_t>((n + nt - 1) / nt, (size_t)ctx->sm_count * per_sm);
        void (*kern)(const uint8_t *, const uint64_t *, const uint8_t *, const uint64_t *, const uint32_t *, size_t,
                     uint32_t, uint32_t *);
        if (bits == 64)
            kern = costs.transpose ? lev_bitpar_tab_kernel<true, 2, uint64_t> : lev_bitpar_tab_kernel<false, 2, uint64_t>;
        else if (bits == 16)
            kern = costs.transpose ? lev_bitpar_tab_kernel<true, 1, uint16_t> : lev_bitpar_tab_kernel<false, 1, uint16_t>;
        else if (planes == 2)
            kern = costs.transpose ? lev_bitpar_tab_kernel<true, 2, uint32_t> : lev_bitpar_tab_kernel<false, 2, uint32_t>;
        else
            kern = costs.transpose ? lev_bitpar_tab_kernel<true, 1, uint32_t> : lev_bitpar_tab_kernel<false, 1, uint32_t>;
        TA_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, nt, smem, st>>>(a, a_off, b, b_off, idx, n, k, out);
        ctx->launches++;
        TA_CUDA(ctx, cudaGetLastError());
        return TA_OK;
    }
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (costs.transpose)
        lev_bitpar32_kernel<true><<<blocks, 128, 0, st>>>(a, a_off, b, b_off, idx, n, k, out);
    else
        lev_bitpar32_kernel<false><<<blocks, 128, 0, st>>>(a, a_off, b, b_off, idx, n, k, out);
    ctx->launches++;
    TA_CUDA(ctx, cudaGetLastError());
    return TA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Search pre-filter: Myers' bit-vector algorithm in its semi-global (search) form, one needle against many
// haystacks.  Computes, for every end position, the unit-cost distance of the best alignment of the whole needle
// ending there (the same cost the reference's scalar search DP produces, src/levenshtein.rs:1723-1806, without the
// match-length bookkeeping) and flags the haystack as soon as one position is within k.  Flagged haystacks are
// then re-run by the exact (cost, length) kernel in search.cu; unflagged ones provably have no match.
//
// Work split: thread (h, s) scans segment s (SEG bytes) of haystack h after a warm-up of 2*N bytes, which is
// enough for the column state to be exact (an optimal alignment of a needle prefix of length i spans at most 2i
// haystack bytes under unit costs).  The 256-entry match-mask table of the needle lives in shared memory.
namespace {

constexpr int FILTER_SEG = TA_SEARCH_SEG;

template <typename W, bool TRANS>
__global__ void __launch_bounds__(128) search_filter_kernel(const uint8_t *__restrict__ needle, uint32_t N,
                                                            const uint8_t *__restrict__ hay,
                                                            const uint64_t *__restrict__ hay_off, size_t n, uint32_t k,
                                                            uint32_t subs, uint32_t *__restrict__ idx_out,
                                                            uint32_t *__restrict__ counter,
                                                            const uint32_t *__restrict__ only_if) {
    if (only_if && *only_if == 0) return;  // queued as the q-gram scan's fallback: runs only if that kernel gave up
    __shared__ W peq[256];
    for (int c = threadIdx.x; c < 256; c += blockDim.x) peq[c] = 0;
    __syncthreads();
    if (threadIdx.x < N) {
        if (sizeof(W) == 4)
            atomicOr((unsigned int *)&peq[needle[threadIdx.x]], 1u << threadIdx.x);
        else
            atomicOr((unsigned long long *)&peq[needle[threadIdx.x]], 1ull << threadIdx.x);
    }
    __syncthreads();

    const size_t h = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    const uint64_t h0 = hay_off[h], h1 = hay_off[h + 1];
    const uint64_t H = h1 - h0;
    const uint64_t seg_begin = (uint64_t)blockIdx.y * FILTER_SEG;
    if (seg_begin >= H) return;
    const uint64_t seg_end = seg_begin + FILTER_SEG < H ? seg_begin + FILTER_SEG : H;
    const uint64_t warm = 2ull * N;
    const uint64_t start = seg_begin > warm ? seg_begin - warm : 0;
    const uint8_t *p = hay + h0;

    W VP = ~(W)0, VN = 0, D0prev = ~(W)0, Eqprev = 0;
    uint32_t score = N;
    const W top = (W)1 << (N - 1);
    // the haystack is streamed in aligned 16-byte vectors and re-aligned in registers (bitpar::Stream)
    bitpar::Stream hs;
    hs.init((intptr_t)(p + start), (uintptr_t)p, (uintptr_t)(p + H - 1));
    const W rowmask = N >= 8 * sizeof(W) ? ~(W)0 : (((W)1 << N) - 1);
    for (uint64_t x0 = start; x0 < seg_end; x0 += 16) {
        uint32_t wds[4];
        hs.take(wds);
        // The score D[N][x] moves by at most one per byte.  While it is more than 16 above k no position of this
        // 16-byte chunk can be a hit: run the bare recurrence and rebuild the score afterwards from the vertical
        // deltas (D[N][x] = #VP - #VN over the needle's rows); otherwise track it byte by byte.
        if (score > k + 16u) {
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const uint32_t ch = (wds[u >> 2] >> (8 * (u & 3))) & 0xffu;
                const W Eq = peq[ch];
                W D0 = (((Eq & VP) + VP) ^ VP) | Eq | VN;
                if (TRANS) {
                    D0 |= ((~D0prev & Eq) << 1) & Eqprev;
                    D0prev = D0;
                    Eqprev = Eq;
                }
                const W HP = (VN | ~(D0 | VP)) << 1;  // row 0 of a search has horizontal delta 0
                const W HN = (D0 & VP) << 1;
                VP = HN | ~(D0 | HP);
                VN = D0 & HP;
            }
            score = sizeof(W) == 4 ? (uint32_t)(__popc((uint32_t)(VP & rowmask)) - __popc((uint32_t)(VN & rowmask)))
                                   : (uint32_t)(__popcll((uint64_t)(VP & rowmask)) - __popcll((uint64_t)(VN & rowmask)));
            continue;
        }
        bool hit = false;
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const uint64_t x = x0 + u;
            const uint32_t ch = (wds[u >> 2] >> (8 * (u & 3))) & 0xffu;
            const W Eq = peq[ch];
            W D0 = (((Eq & VP) + VP) ^ VP) | Eq | VN;
            if (TRANS) {
                D0 |= ((~D0prev & Eq) << 1) & Eqprev;
                D0prev = D0;
                Eqprev = Eq;
            }
            W HP = VN | ~(D0 | VP);
            W HN = D0 & VP;
            score += (HP & top) ? 1u : 0u;
            score -= (HN & top) ? 1u : 0u;
            HP <<= 1;
            HN <<= 1;
            VP = HN | ~(D0 | HP);
            VN = D0 & HP;
            // positions past seg_end read don't-care bytes: they may only raise a flag inside the segment
            hit |= (x >= seg_begin) & (x < seg_end) & (score <= k);
        }
        if (hit) {
            // a match end in [x0, x0 + 16) (rare event): append the codes of the TA_SEARCH_SUB-byte sub-segments from
            // there to the end of this thread's segment, which the exact kernel will re-run
            const uint64_t first = (x0 > seg_begin ? x0 : seg_begin) / TA_SEARCH_SUB, last = (seg_end - 1) / TA_SEARCH_SUB;
            for (uint64_t sub = first; sub <= last; sub++)
                idx_out[atomicAdd(counter, 1u)] = (uint32_t)(h * subs + sub);
            return;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// search_pigeon_kernel: the pre-filter for needles of <= 32 bytes when k + 1 (2k + 1 with transpositions) needle
// pieces are at least PIGEON_MIN_PIECE bytes long.  An alignment with <= k unit-cost edits leaves at least one of
// k + 1 consecutive needle pieces intact (a transposition can damage two, hence 2k + 1), so a haystack position can
// only be a match end if one of the pieces occurs EXACTLY a bounded distance before it.  Exact occurrences of all
// pieces are found with one shift-and recurrence over the needle's match masks -- the very table Myers' recurrence
// uses: D = ((D << 1) | starts) & peq[byte], piece i ends here iff bit (its last needle index) of D is set -- which is
// 2 ALU instructions per haystack byte instead of ~11.  The table is replicated per shared-memory bank ([byte][lane]),
// so the data-dependent look-ups of a warp never conflict.  A candidate (rare: a piece of >= 4 random bytes) is
// verified on the spot by the same thread with Myers' semi-global recurrence over the <= N + 2k bytes that can hold a
// match through this piece occurrence; only confirmed end positions flag their TA_SEARCH_SUB-byte sub-segment, so
// the flags are as precise as the Myers pre-filter's.  Thread (h, s) owns the END positions inside segment s of
// haystack h and scans N + k + (longest piece) bytes before it so that every piece of such a match is seen.
constexpr uint32_t PIGEON_MIN_PIECE = 4;

template <bool TRANS>
__global__ void __launch_bounds__(256) search_pigeon_kernel(const uint8_t *__restrict__ needle, uint32_t N,
                                                            const uint8_t *__restrict__ hay,
                                                            const uint64_t *__restrict__ hay_off, size_t n, uint32_t k,
                                                            uint32_t pieces, uint32_t subs,
                                                            uint32_t *__restrict__ idx_out,
                                                            uint32_t *__restrict__ counter) {
    extern __shared__ uint32_t peqr[];  // [256][32]: entry c of lane l at peqr[c * 32 + l]
    for (uint32_t q = threadIdx.x; q < 256u * 32u; q += blockDim.x) peqr[q] = 0;
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < N * 32u; q += blockDim.x)
        atomicOr(&peqr[(uint32_t)needle[q >> 5] * 32u + (q & 31u)], 1u << (q >> 5));
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t *peq = peqr + lane;  // this lane's copy: entry c at peq[c * 32]

    const size_t h = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n) return;
    const uint64_t h0 = hay_off[h], h1 = hay_off[h + 1];
    const uint64_t H = h1 - h0;
    const uint64_t seg_begin = (uint64_t)blockIdx.y * FILTER_SEG;
    if (seg_begin >= H) return;
    const uint64_t seg_end = seg_begin + FILTER_SEG < H ? seg_begin + FILTER_SEG : H;
    const uint8_t *p = hay + h0;

    // pieces: piece i = needle [s_i, s_i + l_i), l_i = N / pieces (+1 for the first N % pieces)
    const uint32_t base_len = N / pieces, extra = N % pieces;
    uint32_t starts = 0, finals = 0;
    for (uint32_t i = 0, s = 0; i < pieces; i++) {
        const uint32_t l = base_len + (i < extra ? 1u : 0u);
        starts |= 1u << s;
        finals |= 1u << (s + l - 1);
        s += l;
    }
    const uint64_t back = (uint64_t)N + k + base_len + 1;
    uint64_t start = seg_begin > back ? seg_begin - back : 0;
    {  // start on a 16-byte boundary of the haystack's memory when possible: the stream then needs no re-alignment
        const uint64_t mis = (uint64_t)((uintptr_t)(p + start) & 15u);
        if (start >= mis) start -= mis;
    }
    uint32_t flagged = 0;  // sub-segments of this segment already appended (bit = sub index inside the segment)

    // a piece whose last byte is haystack byte q (its last needle index = fin): confirm the match ends it allows
    auto verify = [&](const uint64_t q, const uint32_t fin) {
        const uint32_t r = N - 1 - fin;  // needle bytes after the piece
        // end byte indices of matches through this occurrence: [q + r - k, q + r + k], inside this thread's segment
        uint64_t lo = q + r > k ? q + r - k : 0, hi = q + r + k;
        if (lo < seg_begin) lo = seg_begin;
        if (hi > seg_end - 1) hi = seg_end - 1;
        if (lo > hi) return;
        // such a match starts at most (fin + 1) + k bytes before the piece's end
        const uint64_t st = q + 1 > (uint64_t)fin + 1 + k ? q + 1 - (fin + 1 + k) : 0;
        uint32_t VP = 0xffffffffu, VN = 0, D0prev = 0xffffffffu, Eqprev = 0, score = N;
        const uint32_t top = 1u << (N - 1);
        for (uint64_t t = st; t <= hi; t++) {
            const uint32_t Eq = peq[(uint32_t)p[t] * 32u];
            uint32_t D0 = (((Eq & VP) + VP) ^ VP) | Eq | VN;
            if (TRANS) {
                D0 |= ((~D0prev & Eq) << 1) & Eqprev;
                D0prev = D0;
                Eqprev = Eq;
            }
            uint32_t HP = VN | ~(D0 | VP);
            uint32_t HN = D0 & VP;
            score += (HP & top) ? 1u : 0u;
            score -= (HN & top) ? 1u : 0u;
            HP <<= 1;
            HN <<= 1;
            VP = HN | ~(D0 | HP);
            VN = D0 & HP;
            if (t >= lo && score <= k) {
                const uint32_t sub_in = (uint32_t)((t - seg_begin) / TA_SEARCH_SUB);
                if (!(flagged >> sub_in & 1u)) {
                    flagged |= 1u << sub_in;
                    idx_out[atomicAdd(counter, 1u)] = (uint32_t)(h * subs + t / TA_SEARCH_SUB);
                }
            }
        }
    };

    bitpar::Stream hs;
    hs.init((intptr_t)(p + start), (uintptr_t)p, (uintptr_t)(p + H - 1));
    uint32_t D = 0;
    for (uint64_t x0 = start; x0 < seg_end; x0 += 16) {
        uint32_t wds[4];
        hs.take(wds);
        const uint32_t Dstart = D;
        uint32_t seen = 0;
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const uint32_t ch = bitpar::byte_of(wds[u >> 2], u & 3);
            D = ((D << 1) | starts) & peq[ch * 32u];
            seen |= D;
        }
        if (seen & finals) {  // some piece ends inside this chunk: replay it byte by byte
            uint32_t Dr = Dstart;
            for (int u = 0; u < 16; u++) {
                const uint64_t q = x0 + (uint64_t)u;
                if (q >= H) break;  // bytes past the haystack are don't-care padding of the last vector
                const uint32_t ch = wds[0] & 0xffu;  // (the words are shifted along: no dynamic register indexing)
#pragma unroll
                for (int w = 0; w < 3; w++) wds[w] = bitpar::funnel_r(wds[w], wds[w + 1], 8);
                wds[3] >>= 8;
                Dr = ((Dr << 1) | starts) & peq[ch * 32u];
                uint32_t f = Dr & finals;
                while (f) {
                    const uint32_t fin = (uint32_t)__ffs((int)f) - 1u;
                    f &= f - 1u;
                    verify(q, fin);
                }
            }
        }
    }
}


// search_pigeon_staged_kernel: the same filter with the haystack STAGED THROUGH SHARED MEMORY.  In the kernel above
// every lane streams its own segment, so one warp-wide 16-byte load touches 32 different 128-byte lines and the LSU
// data pipe, not arithmetic, bounds the kernel.  Here a warp owns 32 consecutive (haystack, segment) items and moves
// them in steps of 64 bytes per item: the 32 x 64-byte rows of a step are fetched by coalesced 16-byte cp.async
// (LDGSTS, four lanes per row, no register staging) into a padded per-warp tile, double-buffered so that the copy
// of step t + 1 overlaps the arithmetic of step t, and every lane then reads its own row back with four
// conflict-free LDS.128.  Rows start on the 16-byte boundary at or below the item's first byte; vectors past the
// haystack's last one are clamped like bitpar::Stream does (don't-care bytes, never flagged: see verify()).
constexpr int PIGEON_ROW = 80;  // 64 data bytes + 16 of padding: rows of a quarter-warp fall into distinct banks

template <bool TRANS>
__global__ void __launch_bounds__(256) search_pigeon_staged_kernel(const uint8_t *__restrict__ needle, uint32_t N,
                                                                   const uint8_t *__restrict__ hay,
                                                                   const uint64_t *__restrict__ hay_off, size_t n,
                                                                   uint32_t k, uint32_t pieces, uint32_t subs,
                                                                   uint32_t segs, uint32_t *__restrict__ idx_out,
                                                                   uint32_t *__restrict__ counter,
                                                                   const uint32_t *__restrict__ only_if) {
    // queued behind search_qgram_kernel as its fallback: runs only if that kernel gave up (flag set on the device)
    if (only_if && *only_if == 0) return;
    extern __shared__ __align__(16) uint8_t pg_smem[];
    uint32_t *peqr = (uint32_t *)pg_smem;  // [256][32]
    for (uint32_t q = threadIdx.x; q < 256u * 32u; q += blockDim.x) peqr[q] = 0;
    __syncthreads();
    for (uint32_t q = threadIdx.x; q < N * 32u; q += blockDim.x)
        atomicOr(&peqr[(uint32_t)needle[q >> 5] * 32u + (q & 31u)], 1u << (q >> 5));
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t *peq = peqr + lane;
    uint8_t *tile = pg_smem + 256 * 32 * 4 + (size_t)warp * (2 * 32 * PIGEON_ROW);

    const size_t item = ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32 + lane;
    bool active = item < n * (size_t)segs;
    const size_t h = active ? item / segs : 0;
    const uint32_t sidx = active ? (uint32_t)(item % segs) : 0;
    const uint64_t h0 = hay_off[h], h1 = hay_off[h + 1];
    const uint64_t H = h1 - h0;
    const uint64_t seg_begin = (uint64_t)sidx * FILTER_SEG;
    if (seg_begin >= H) active = false;
    const uint64_t seg_end = seg_begin + FILTER_SEG < H ? seg_begin + FILTER_SEG : H;
    const uint8_t *p = hay + h0;

    const uint32_t base_len = N / pieces, extra = N % pieces;
    uint32_t starts = 0, finals = 0;
    for (uint32_t i = 0, s = 0; i < pieces; i++) {
        const uint32_t l = base_len + (i < extra ? 1u : 0u);
        starts |= 1u << s;
        finals |= 1u << (s + l - 1);
        s += l;
    }
    const uint64_t back = (uint64_t)N + k + base_len + 1;
    const uint64_t start = seg_begin > back ? seg_begin - back : 0;
    // row geometry: a0 = 16-byte boundary at or below the first scanned byte, x0 = its byte index in the haystack
    const uintptr_t a0 = active ? ((uintptr_t)(p + start) & ~(uintptr_t)15) : 0;
    const int64_t x0 = (int64_t)a0 - (int64_t)(uintptr_t)p;
    const int vlast = active ? (int)((((uintptr_t)(p + H - 1) & ~(uintptr_t)15) - a0) >> 4) : -1;
    const int own_steps = active ? (int)(((int64_t)seg_end - x0 + 63) >> 6) : 0;
    int nsteps = own_steps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nsteps = max(nsteps, __shfl_xor_sync(0xffffffffu, nsteps, o));

    // loader duty: this lane fetches vector slot (lane & 3) of rows (lane >> 2) + 8 r, r = 0..3, of every step
    const uint32_t slot = lane & 3u;
    unsigned long long lbase[4];
    int lvlast[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int jr = (int)(lane >> 2) + 8 * r;
        lbase[r] = __shfl_sync(0xffffffffu, (unsigned long long)a0, jr);
        lvlast[r] = __shfl_sync(0xffffffffu, vlast, jr);
    }
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
    auto issue = [&](const int t, const int buf) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (lvlast[r] >= 0) {
                const int jr = (int)(lane >> 2) + 8 * r;
                int v = 4 * t + (int)slot;
                v = v < lvlast[r] ? v : lvlast[r];
                const unsigned long long src = lbase[r] + 16ull * (unsigned long long)v;
                const uint32_t dst = tile_s + (uint32_t)(buf * 32 * PIGEON_ROW + jr * PIGEON_ROW) + 16u * slot;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    uint32_t flagged = 0;
    auto verify = [&](const uint64_t q, const uint32_t fin) {
        const uint32_t r = N - 1 - fin;
        uint64_t lo = q + r > k ? q + r - k : 0, hi = q + r + k;
        if (lo < seg_begin) lo = seg_begin;
        if (hi > seg_end - 1) hi = seg_end - 1;
        if (lo > hi) return;
        const uint64_t st = q + 1 > (uint64_t)fin + 1 + k ? q + 1 - (fin + 1 + k) : 0;
        uint32_t VP = 0xffffffffu, VN = 0, D0prev = 0xffffffffu, Eqprev = 0, score = N;
        const uint32_t top = 1u << (N - 1);
        for (uint64_t t = st; t <= hi; t++) {
            const uint32_t Eq = peq[(uint32_t)p[t] * 32u];
            uint32_t D0 = (((Eq & VP) + VP) ^ VP) | Eq | VN;
            if (TRANS) {
                D0 |= ((~D0prev & Eq) << 1) & Eqprev;
                D0prev = D0;
                Eqprev = Eq;
            }
            uint32_t HP = VN | ~(D0 | VP);
            uint32_t HN = D0 & VP;
            score += (HP & top) ? 1u : 0u;
            score -= (HN & top) ? 1u : 0u;
            HP <<= 1;
            HN <<= 1;
            VP = HN | ~(D0 | HP);
            VN = D0 & HP;
            if (t >= lo && score <= k) {
                const uint32_t sub_in = (uint32_t)((t - seg_begin) / TA_SEARCH_SUB);
                if (!(flagged >> sub_in & 1u)) {
                    flagged |= 1u << sub_in;
                    idx_out[atomicAdd(counter, 1u)] = (uint32_t)(h * subs + t / TA_SEARCH_SUB);
                }
            }
        }
    };

    uint32_t D = 0;
    if (nsteps > 0) issue(0, 0);
    for (int t = 0; t < nsteps; t++) {
        if (t + 1 < nsteps) {
            issue(t + 1, (t + 1) & 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        if (t < own_steps) {
            const uint4 *row = (const uint4 *)(tile + (t & 1) * 32 * PIGEON_ROW + lane * PIGEON_ROW);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int64_t xc = x0 + 64 * (int64_t)t + 16 * c;  // haystack index of the chunk's first byte
                if (xc >= (int64_t)seg_end) break;
                const uint4 v4 = row[c];
                uint32_t wds[4] = {v4.x, v4.y, v4.z, v4.w};
                const uint32_t Dstart = D;
                uint32_t seen = 0;
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const uint32_t ch = bitpar::byte_of(wds[u >> 2], u & 3);
                    D = ((D << 1) | starts) & peq[ch * 32u];
                    seen |= D;
                }
                if (seen & finals) {  // some piece ends inside this chunk: replay it byte by byte
                    uint32_t Dr = Dstart;
                    for (int u = 0; u < 16; u++) {
                        const int64_t q = xc + u;
                        if (q >= (int64_t)H) break;  // don't-care padding of the last vector
                        const uint32_t ch = wds[0] & 0xffu;
#pragma unroll
                        for (int w = 0; w < 3; w++) wds[w] = bitpar::funnel_r(wds[w], wds[w + 1], 8);
                        wds[3] >>= 8;
                        Dr = ((Dr << 1) | starts) & peq[ch * 32u];
                        uint32_t f = q >= 0 ? (Dr & finals) : 0u;  // bytes below the haystack's first are padding too
                        while (f) {
                            const uint32_t fin = (uint32_t)__ffs((int)f) - 1u;
                            f &= f - 1u;
                            verify((uint64_t)q, fin);
                        }
                    }
                }
            }
        }
        __syncwarp();  // everyone is done with buffer t & 1 before step t + 2 overwrites it
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// search_qgram_kernel + search_qgram_resolve_kernel (round 2): the exact-piece filter at memory speed, for pieces of
// >= 7 bytes.  The shift-and scan above looks at EVERY haystack byte (8.3 instructions per byte: the kernel is
// issue-bound at a third of the HBM rate).  But a piece of l >= 7 bytes that occurs exactly in the haystack covers at
// least one 4-byte-ALIGNED word of memory completely, wherever it lies, and that word is one of the l - 3 four-byte
// substrings ("4-grams") of the piece.  So it is enough to look at the aligned words: hash each one into a 256 Kbit
// shared-memory bitmap of the needle's 4-grams (all pieces, all offsets: a few dozen entries), ~2 instructions per
// haystack byte, with the haystacks streamed as ONE flat byte range by coalesced 16-byte loads -- no per-haystack work
// items, no staging.  A bitmap hit (a real piece word, or one false positive per ~13 000 words) is compared with the
// 4-gram list; an exact match is QUEUED (flat offset + word) and the scan goes on.  The resolve kernel takes one queue
// entry per thread: finds the haystack, compares the whole piece, and runs the same bounded Myers verification as above
// over the match ends this occurrence allows; confirmed ends flag their 128-byte sub-segment exactly once (a global
// bitmap with atomicOr de-duplicates across threads and pieces).  (Resolving inline was the first form: a planted
// needle gives ~7 gram hits in one or two lanes of ONE warp, ~10 us of dependent loads each, and the scan waited 70 us
// for the slowest warps -- SMSP active cycles min 121 k / avg 193 k / max 342 k.)
// Low-entropy input (DNA, text: common 4-grams everywhere) would drown the queue; when it overflows the scan sets a
// flag and stops, the resolve kernel does nothing, and the shift-and kernel queued behind them (`only_if`) does the
// job instead -- same output contract, no host round trip.
// Coverage: an alignment with <= k edits leaves one of the k + 1 (2k + 1) pieces intact; the intact piece contains an
// aligned word; that word is in the bitmap and the list; its verification window holds the match.
struct QGram {
    uint32_t gram;  // the four bytes, little endian as loaded from memory
    uint8_t off;    // offset of the word in its piece
    uint8_t len;    // piece length
    uint16_t fin;   // needle index of the piece's last byte
};
struct QCand {
    uint64_t g;  // flat byte offset of the word (from `hay`)
    uint32_t word;
    uint32_t pad;
};
constexpr uint32_t QG_LOG = 18;  // 256 Kbit = 32 KB of shared memory
constexpr uint32_t QG_BITS = 1u << QG_LOG;
constexpr int QG_MAX = 64;
constexpr uint32_t QG_CHUNK = 512;  // vectors per chunk: 8 KB
constexpr uint32_t QG_REGIONS = 8;  // chunk counters (one 128-byte line each)
constexpr uint32_t QG_GRAN = 16;    // bytes of end positions per flagged granule

__device__ __forceinline__ uint32_t qg_mul(uint32_t x) { return x * 0x9E3779B1u; }  // hash = top QG_LOG bits

// the chunk counter (see the note at its use in the scan loop)
__device__ __forceinline__ uint32_t qg_next_chunk(uint32_t *ctr) {
    uint32_t r;
    asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(r) : "l"(ctr) : "memory");
    return r;
}

// the 4-grams of every piece, in a fixed order (both kernels build the same list); returns their number
__device__ __forceinline__ uint32_t qg_build(const uint8_t *nd, uint32_t N, uint32_t pieces, QGram *grams) {
    const uint32_t base_len = N / pieces, extra = N % pieces;
    uint32_t cnt = 0;
    for (uint32_t i = 0, s0 = 0; i < pieces; i++) {
        const uint32_t l = base_len + (i < extra ? 1u : 0u);
        for (uint32_t o = 0; o + 4 <= l && cnt < (uint32_t)QG_MAX; o++) {
            grams[cnt].gram = (uint32_t)nd[s0 + o] | ((uint32_t)nd[s0 + o + 1] << 8) | ((uint32_t)nd[s0 + o + 2] << 16) |
                              ((uint32_t)nd[s0 + o + 3] << 24);
            grams[cnt].off = (uint8_t)o, grams[cnt].len = (uint8_t)l, grams[cnt].fin = (uint16_t)(s0 + l - 1);
            cnt++;
        }
        s0 += l;
    }
    return cnt;
}

struct QGramShared {
    uint32_t bitmap[QG_BITS / 32];
    QGram grams[QG_MAX];
    uint32_t n_grams;
    uint8_t needle[64];
};

// a word that hit the bitmap: if it really is one of the 4-grams, queue it.  Out of line: it runs for one word in ~10^4
// and the scan loop around it must stay small (inlined 16 times it made the kernel 110 KB of SASS, every probe its own
// reconvergence region, and the scan ran at a quarter of the issue rate with `no_instruction` / `wait` stalls).
// Returns false when the queue is full (the caller stops scanning).
__device__ __noinline__ bool qgram_push(const QGramShared &sh, const uint32_t word, const uint64_t g, QCand *__restrict__ queue,
                                        uint32_t *__restrict__ qcount, const uint32_t qcap, uint32_t *__restrict__ gave_up) {
    const uint32_t ng = sh.n_grams;
    bool any = false;
    for (uint32_t j = 0; j < ng; j++) any |= sh.grams[j].gram == word;
    if (!any) return true;
    const uint32_t slot = atomicAdd(qcount, 1u);
    if (slot >= qcap) {
        *gave_up = 1u;
        return false;
    }
    QCand c;
    c.g = g, c.word = word, c.pad = 0;
    queue[slot] = c;
    return true;
}

template <int U, int OCC>  // U 16-byte loads per thread and round, OCC resident CTAs per SM
__global__ void __launch_bounds__(256, OCC) search_qgram_kernel(const uint8_t *__restrict__ needle, uint32_t N,
                                                              const uint8_t *__restrict__ hay,
                                                              const uint64_t *__restrict__ hay_off, size_t n,
                                                              uint32_t pieces, QCand *__restrict__ queue,
                                                              uint32_t *__restrict__ qcount, uint32_t qcap,
                                                              uint32_t *__restrict__ gave_up,
                                                              uint32_t *__restrict__ chunk_ctr) {
    __shared__ QGramShared sh;
    for (uint32_t q = threadIdx.x; q < QG_BITS / 32; q += blockDim.x) sh.bitmap[q] = 0;
    if (threadIdx.x < 64) sh.needle[threadIdx.x] = threadIdx.x < N ? needle[threadIdx.x] : 0;
    __syncthreads();
    if (threadIdx.x == 0) {  // a few dozen entries: one thread, needle bytes from shared memory
        const uint32_t cnt = qg_build(sh.needle, N, pieces, sh.grams);
        for (uint32_t j = 0; j < cnt; j++) {
            const uint32_t hsh = qg_mul(sh.grams[j].gram) >> (32 - QG_LOG);
            sh.bitmap[hsh >> 5] |= 1u << (hsh & 31u);
        }
        sh.n_grams = cnt;
    }
    __syncthreads();
    if (n == 0) return;
    const uint64_t B0 = hay_off[0], B1 = hay_off[n];
    if (B1 - B0 < 4) return;
    // aligned words that lie fully inside the haystack bytes: word index j = address / 4, [first_w, end_w)
    const uintptr_t abase = (uintptr_t)hay;
    const uint64_t first_w = (abase + B0 + 3) >> 2, end_w = (abase + B1) >> 2;

    // bit 0 of the result = the word is in the bitmap (branch-free: ~7 instructions, independent across words)
    auto test = [&](const uint32_t word) {
        const uint32_t m = qg_mul(word);
        return sh.bitmap[m >> (32 - QG_LOG + 5)] >> ((m >> (32 - QG_LOG)) & 31u);
    };
    auto ldw = [](uint64_t wi) { return __ldg(reinterpret_cast<const uint32_t *>(wi << 2)); };  // wi = address / 4
    auto ldv = [](uint64_t v) { return __ldg(reinterpret_cast<const uint4 *>(v << 4)); };       // v = address / 16
    auto slow = [&](const uint64_t wi) {  // one word, rare: test again and follow up
        const uint32_t word = ldw(wi);
        if (test(word) & 1u) return qgram_push(sh, word, (wi << 2) - abase, queue, qcount, qcap, gave_up);
        return true;
    };

    // body: 16-byte vectors [v0, v1), each four aligned words; the (at most three) words before and after it one by one.
    // The vectors are handed out in chunks of QG_CHUNK (8 KB) through a device counter, a warp at a time: with a fixed
    // share per warp the scan ran at ~5.4 TB/s until the fastest warps were done and then waited as long again for the
    // slowest ones (SM active cycles min 151 k / max 282 k).  A warp asks for its next chunk before it works on the
    // current one, so the counter's round trip is never waited for.  Within a round U independent 16-byte
    // loads are in flight per thread and their 16 bitmap tests are OR-ed into one flag: one branch per 64 bytes.
    const uint64_t v0 = (first_w + 3) >> 2, v1 = end_w >> 2;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool go = true;
    if (v0 >= v1) {  // no whole vector: a handful of words
        if (tid == 0)
            for (uint64_t wi = first_w; go && wi < end_w; wi++) go = slow(wi);
        return;
    }
    if (tid < 8) {
        for (uint64_t wi = first_w + tid; go && wi < 4 * v0; wi += 8) go = slow(wi);
        for (uint64_t wi = 4 * v1 + tid; go && wi < end_w; wi += 8) go = slow(wi);
    }
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t n_chunks = (v1 - v0 + QG_CHUNK - 1) / QG_CHUNK;
    constexpr uint32_t ROUND = 32 * U, ROUNDS = QG_CHUNK / ROUND;  // vectors per round of the warp, rounds per chunk
    // Chunk hand-out: the chunks are cut into QG_REGIONS contiguous regions, each with its own counter in its own
    // 128-byte line; a warp works its way through region (warp id mod QG_REGIONS) and, when that is used up, through
    // the others.  (ONE counter for all 50 k chunks was the bottleneck of the scan: an L2 atomic unit serves one address
    // at ~2 ns per request, whatever the number of warps -- 2 to 6 CTAs per SM all ran at 110 us, a third of the stall
    // samples on the counter's reply.)
    const uint64_t cpr = (n_chunks + QG_REGIONS - 1) / QG_REGIONS;  // chunks per region
    uint32_t region = (uint32_t)((tid >> 5) % QG_REGIONS), tried = 0;
    auto region_size = [&](uint32_t r) -> uint64_t {
        const uint64_t lo = (uint64_t)r * cpr;
        return lo >= n_chunks ? 0 : (n_chunks - lo < cpr ? n_chunks - lo : cpr);
    };
    auto request = [&]() -> uint32_t { return lane == 0 ? qg_next_chunk(chunk_ctr + region * 32u) : 0u; };
    // ticket (lane 0) -> chunk index, or n_chunks when everything has been handed out; moves on to the next region
    // (one exposed round trip each) when the ticket is past the end of the current one
    auto resolve = [&](uint32_t ticket) -> uint64_t {
        uint32_t li = __shfl_sync(0xffffffffu, ticket, 0);
        while ((uint64_t)li >= region_size(region)) {
            if (++tried == QG_REGIONS) return n_chunks;
            region = region + 1 == QG_REGIONS ? 0 : region + 1;
            li = __shfl_sync(0xffffffffu, request(), 0);
        }
        return (uint64_t)region * cpr + li;
    };
    auto full_chunk = [&](uint64_t cc) { return v0 + (cc + 1) * QG_CHUNK <= v1; };
    auto partial = [&](uint64_t cc) {  // the last chunk when it is not whole: word by word
        for (uint64_t v = v0 + cc * QG_CHUNK + lane; go && v < v1; v += 32)
#pragma unroll 1
            for (int i = 0; i < 4; i++) go = go && slow(4 * v + i);
    };
    // software pipeline over rounds, across chunk boundaries: the loads of round r + 1 are issued before round r is
    // tested, so a warp always has U to 2U loads in flight (with load -> test -> load the scan was latency-bound at
    // 4.2 TB/s: most warps waiting on their loads, issue slots 35 % busy)
    uint64_t c = resolve(request());
    while (c < n_chunks && !full_chunk(c)) {  // drew the partial last chunk first: do it and draw again
        partial(c);
        c = resolve(request());
    }
    if (c >= n_chunks) return;
    uint32_t ticket = request();
    uint32_t it = 0;
    uint64_t vcur = v0 + c * QG_CHUNK + lane;
    uint4 x[U];
#pragma unroll
    for (int u = 0; u < U; u++) x[u] = ldv(vcur + u * 32);
    while (true) {
        // where the next round is
        bool have_next = true, new_chunk = false;
        if (++it == ROUNDS) {
            it = 0;
            c = resolve(ticket);
            while (c < n_chunks && !full_chunk(c)) {
                partial(c);
                c = resolve(request());
            }
            have_next = c < n_chunks;
            new_chunk = have_next;
        }
        const uint64_t vnext = v0 + c * QG_CHUNK + it * ROUND + lane;
        uint4 y[U];
        if (have_next) {
#pragma unroll
            for (int u = 0; u < U; u++) y[u] = ldv(vnext + u * 32);
        }
        // the request for the chunk after this one goes out AFTER the loads: ptxas turns a one-lane atom.add (atomicAdd()
        // or inline PTX, atom.inc too) into its warp-aggregation sequence -- elect, ATOMG, SHFL of the returned value --
        // and that SHFL waits for the round trip on the spot
        if (new_chunk) ticket = request();
        uint32_t any = 0;
#pragma unroll
        for (int u = 0; u < U; u++) any |= test(x[u].x) | test(x[u].y) | test(x[u].z) | test(x[u].w);
        if (any & 1u) {
#pragma unroll 1
            for (int u = 0; u < U; u++)
#pragma unroll 1
                for (int i = 0; i < 4; i++) go = go && slow(4 * (vcur + u * 32) + i);
        }
        if (__any_sync(0xffffffffu, !go)) return;  // the queue is full: the fallback kernel takes over
        if (!have_next) break;
#pragma unroll
        for (int u = 0; u < U; u++) x[u] = y[u];
        vcur = vnext;
    }
}

// one queued word per thread: which haystack, which pieces, and what the occurrence allows
template <typename W, bool TRANS>  // W: the verification's bit-vector word (uint32_t for needles <= 32, else uint64_t)
__global__ void __launch_bounds__(128) search_qgram_resolve_kernel(const uint8_t *__restrict__ needle, uint32_t N,
                                                                   const uint8_t *__restrict__ hay,
                                                                   const uint64_t *__restrict__ hay_off, size_t n,
                                                                   uint32_t k, uint32_t pieces, uint32_t subs,
                                                                   uint32_t gran, const QCand *__restrict__ queue,
                                                                   const uint32_t *__restrict__ qcount,
                                                                   const uint32_t *__restrict__ gave_up,
                                                                   uint32_t *__restrict__ sub_flags,
                                                                   uint32_t *__restrict__ idx_out,
                                                                   uint32_t *__restrict__ counter) {
    if (*gave_up) return;
    if ((uint64_t)blockIdx.x * blockDim.x >= *qcount) return;  // the grid is sized for a full queue; most blocks have nothing
    __shared__ W peq[256];
    __shared__ QGram grams[QG_MAX];
    __shared__ uint8_t nd[64];
    __shared__ uint32_t n_grams;
    for (uint32_t q = threadIdx.x; q < 256; q += blockDim.x) peq[q] = 0;
    if (threadIdx.x < 64) nd[threadIdx.x] = threadIdx.x < N ? needle[threadIdx.x] : 0;
    __syncthreads();
    if (threadIdx.x < N) {
        if (sizeof(W) == 4)
            atomicOr((unsigned int *)&peq[nd[threadIdx.x]], 1u << threadIdx.x);
        else
            atomicOr((unsigned long long *)&peq[nd[threadIdx.x]], 1ull << threadIdx.x);
    }
    if (threadIdx.x == 0) n_grams = qg_build(nd, N, pieces, grams);
    __syncthreads();
    const uint32_t total = *qcount, ng = n_grams;
    const uint64_t B0 = hay_off[0], span = hay_off[n] - B0;
    const uint64_t avg_len = span / n ? span / n : 1;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const uint64_t g = queue[e].g;
        const uint32_t word = queue[e].word;
        // the haystack that holds byte g (largest h with hay_off[h] <= g): interpolate, then bisect what is left
        size_t lo = 0, hi = n;
        {
            const size_t est = (size_t)((g - B0) / avg_len);  // exact for equal-length haystacks
            const size_t e0 = est < n ? est : n - 1;
            if (hay_off[e0] <= g) {
                lo = e0;
                if (hay_off[e0 + 1] > g) hi = e0 + 1;
            } else {
                hi = e0;
            }
        }
        while (hi - lo > 1) {
            const size_t mid = lo + (hi - lo) / 2;
            if (hay_off[mid] <= g)
                lo = mid;
            else
                hi = mid;
        }
        const size_t h = lo;
        const uint64_t h0 = hay_off[h], H = hay_off[h + 1] - h0;
        const uint8_t *p = hay + h0;
        const uint64_t x = g - h0;
        // Find the matching 4-gram FIRST and do the work after the search loop: with the work inside `for j`, lanes whose
        // words are different grams ran their verifications one after the other (ncu: 1 active thread per instruction,
        // 128 us for 2 400 entries).  A word can equal several grams only if the needle repeats itself: outer loop.
        for (uint32_t j = 0;; j++) {
            while (j < ng && grams[j].gram != word) j++;
            if (j >= ng) break;
            const uint32_t off = grams[j].off, len = grams[j].len, fin = grams[j].fin;
            if (x < off) continue;            // the piece would start before the haystack
            const uint64_t q = x - off + len - 1;  // haystack index of the piece's last byte
            if (q >= H) continue;             // ... or run past it (the word straddles two haystacks)
            uint32_t diff = 0;                // the rest of the piece
            for (uint32_t i = 0; i < len; i++)
                if (i < off || i >= off + 4) diff |= (uint32_t)p[x - off + i] ^ (uint32_t)nd[fin + 1 - len + i];
            if (diff) continue;
            const uint32_t r = N - 1 - fin;  // needle bytes after the piece
            // end byte indices of matches through this occurrence: [q + r - k, q + r + k]
            uint64_t elo = q + r > k ? q + r - k : 0, ehi = q + r + k;
            if (ehi > H - 1) ehi = H - 1;
            if (elo > ehi) continue;
            const uint64_t st = q + 1 > (uint64_t)fin + 1 + k ? q + 1 - (fin + 1 + k) : 0;
            W VP = ~(W)0, VN = 0, D0prev = ~(W)0, Eqprev = 0;
            uint32_t score = N, hit_subs = 0;
            const uint64_t sub0 = elo / gran;
            const W top = (W)1 << (N - 1);
            for (uint64_t t = st; t <= ehi; t++) {
                const W Eq = peq[p[t]];
                W D0 = (((Eq & VP) + VP) ^ VP) | Eq | VN;
                if (TRANS) {
                    D0 |= ((~D0prev & Eq) << 1) & Eqprev;
                    D0prev = D0;
                    Eqprev = Eq;
                }
                W HP = VN | ~(D0 | VP);
                W HN = D0 & VP;
                score += (HP & top) ? 1u : 0u;
                score -= (HN & top) ? 1u : 0u;
                HP <<= 1;
                HN <<= 1;
                VP = HN | ~(D0 | HP);
                VN = D0 & HP;
                // ends elo .. ehi span a few granules (2k + 1 <= 19 positions): note which, flag after the loop
                if (t >= elo && score <= k) hit_subs |= 1u << (uint32_t)(t / gran - sub0);
            }
            for (uint32_t w = 0; hit_subs >> w; w++) {
                if (!(hit_subs >> w & 1u)) continue;
                const uint64_t code = (uint64_t)h * subs + sub0 + w;
                const uint32_t bit = 1u << (code & 31u);
                if (!(atomicOr(&sub_flags[code >> 5], bit) & bit)) idx_out[atomicAdd(counter, 1u)] = (uint32_t)code;
            }
        }
    }
}

}  // namespace

// appends the codes (haystack * subs + sub-segment, unordered) of the TA_SEARCH_SUB-byte haystack sub-segments that
// contain at least one end position of unit-cost distance <= k to idx_out, counting them in *counter (must be zero on
// entry); *subs_out = sub-segments per haystack.  Needs needle_len in [1, 64], n * subs < 2^32; max_hay = longest
// haystack.  Needles of <= 32 bytes whose k + 1 pieces are long enough use the exact-piece filter (search_pigeon_kernel),
// the rest Myers' recurrence over the whole haystack (search_filter_kernel); TA_SEARCH_FILTER=myers|pigeon forces one.
int ta_launch_search_filter(ta_ctx *ctx, const uint8_t *needle_dev, uint32_t needle_len, const uint8_t *hay,
                            const uint64_t *hay_off, size_t n, uint64_t max_hay, uint32_t k, bool transpose,
                            uint32_t *idx_out, uint32_t *counter, ta_filter_out *fo, cudaStream_t st) {
    if (needle_len == 0 || needle_len > 64) return TA_ERR_TOO_LARGE;
    const uint64_t subs = max_hay ? (max_hay + TA_SEARCH_SUB - 1) / TA_SEARCH_SUB : 1;
    const uint64_t segs = max_hay ? (max_hay + FILTER_SEG - 1) / FILTER_SEG : 1;
    *fo = ta_filter_out();
    fo->segs = (uint32_t)subs, fo->gran = TA_SEARCH_SUB;
    if (n == 0 || max_hay == 0) return TA_OK;
    if (segs > 65535 || (uint64_t)n * subs > 0xFFFFFFF0ull) return TA_ERR_TOO_LARGE;
    const dim3 grid((unsigned)((n + 127) / 128), (unsigned)segs);
    static const char *force = getenv("TA_SEARCH_FILTER");
    const uint32_t pieces = transpose ? 2 * k + 1 : k + 1;
    bool pigeon = needle_len <= 32 && pieces <= needle_len && needle_len / pieces >= PIGEON_MIN_PIECE;
    if (force && force[0] == 'm') pigeon = false;
    if (force && force[0] == 'p' && needle_len <= 32 && pieces <= needle_len) pigeon = true;
    // pieces of >= 7 bytes: aligned-word sampling at memory speed (TA_SEARCH_FILTER=pigeon|myers keep the others testable)
    bool qgram = needle_len <= 64 && pieces <= needle_len && needle_len / pieces >= 7 &&
                 pieces * (needle_len / pieces + 1 - 3) <= (uint32_t)QG_MAX;
    if (force && force[0] != 'q') qgram = false;
    // Three launches more than the scanning filters: pays from ~64 MB of haystacks per call (needles <= 32: 0.45 us/MB
    // for the shift-and kernel against 0.2 us/MB + ~20 us; 8-way strong scaling of cfg 4 leaves 51 MB per GPU and was
    // 0.161 ms with the shift-and kernel, 0.174 ms with the scan) or ~8 MB (longer needles: the Myers filter takes
    // 2.3 us/MB).  TA_SEARCH_FILTER=qgram takes it at any size (tests).
    if (!(force && force[0] == 'q') && (uint64_t)n * max_hay < (needle_len <= 32 ? (64ull << 20) : (8ull << 20))) qgram = false;
    const uint32_t *only_if = nullptr;  // set when the shift-and kernel below is only the q-gram scan's fallback
    if (qgram) {
        // the resolve kernel confirms END POSITIONS, so it can flag much finer granules than the scanning filters: 16
        // bytes instead of TA_SEARCH_SUB -- the exact kernel's serial chain per item is margin + granule + 31 steps
        const uint64_t subs16 = (max_hay + QG_GRAN - 1) / QG_GRAN;
        const bool fine = (uint64_t)n * subs16 <= 0xFFFFFFF0ull;
        const uint64_t qsubs = fine ? subs16 : subs;
        const uint32_t qgran = fine ? QG_GRAN : (uint32_t)TA_SEARCH_SUB;
        const uint64_t nbits = (uint64_t)n * qsubs;
        const size_t words = (size_t)((nbits + 31) / 32);
        // queue: one entry per exact 4-gram match; ~7 per real occurrence of a 32-byte needle on high-entropy input, so
        // 16 per 4 KB leave room for a match in every haystack.  More than that (one per 64 haystack words) means common
        // 4-grams (DNA: ~80 per 4 KB): give up and let the shift-and kernel do it.
        static const long env_cap = getenv("TA_QGRAM_QCAP") ? atol(getenv("TA_QGRAM_QCAP")) : 0;  // testing: force the fallback
        const uint64_t total_bytes = (uint64_t)n * max_hay;  // upper bound of the flat range
        const uint32_t qcap = env_cap > 0 ? (uint32_t)env_cap : (uint32_t)std::min<uint64_t>(total_bytes / TA_QGRAM_BYTES_PER_ENTRY + 4096, 1u << 24);
        // workspace: [sub-segment flags | chunk counters, one 128-byte line each | queue]; one memset clears the first two
        const size_t flag_bytes = (words * sizeof(uint32_t) + 127) & ~(size_t)127;
        const size_t ctr_bytes = (size_t)QG_REGIONS * 128;
        int rc = ta_dev_reserve(ctx, ctx->d_work[2], flag_bytes + ctr_bytes + (size_t)qcap * sizeof(QCand));
        if (rc != TA_OK) return rc;
        uint32_t *sub_flags = (uint32_t *)ctx->d_work[2].p;
        uint32_t *chunk_ctr = (uint32_t *)((uint8_t *)ctx->d_work[2].p + flag_bytes);
        QCand *queue = (QCand *)((uint8_t *)ctx->d_work[2].p + flag_bytes + ctr_bytes);
        uint32_t *qcount = ctx->d_flags + 6, *gave_up = ctx->d_flags + 7;  // zeroed by search_device with its own counters
        TA_CUDA(ctx, cudaMemsetAsync(sub_flags, 0, flag_bytes + ctr_bytes, st));
        // one wave of resident CTAs, chunks handed out dynamically.  Four loads per thread and round, 4 CTAs/SM (64
        // registers): measured against 2 loads x 5 CTAs, 8 x 3 and 4 x 5 (spills) -- 0.302 / 0.308 / 0.315 / 0.346 ms
        // for the whole search step; TA_QGRAM_CTAS changes the grid (2 and 3 CTAs/SM: 0.316 / 0.310 ms)
        static const int env_ctas = getenv("TA_QGRAM_CTAS") ? atoi(getenv("TA_QGRAM_CTAS")) : 0;
        const unsigned blocks = (unsigned)ctx->sm_count * (unsigned)(env_ctas > 0 && env_ctas <= 4 ? env_ctas : 4);
        search_qgram_kernel<4, 4><<<blocks, 256, 0, st>>>(needle_dev, needle_len, hay, hay_off, n, pieces, queue, qcount, qcap, gave_up, chunk_ctr);
        ctx->launches++;
        TA_CUDA(ctx, cudaGetLastError());
        const unsigned rblocks = (unsigned)std::min<uint64_t>(((uint64_t)qcap + 127) / 128, (uint64_t)ctx->sm_count * 8);
#define TA_RESOLVE(W, T)                                                                                                      \
    search_qgram_resolve_kernel<W, T><<<rblocks, 128, 0, st>>>(needle_dev, needle_len, hay, hay_off, n, k, pieces, (uint32_t)qsubs, \
                                                               qgran, queue, qcount, gave_up, sub_flags, idx_out, counter)
        if (needle_len <= 32) {
            if (transpose)
                TA_RESOLVE(uint32_t, true);
            else
                TA_RESOLVE(uint32_t, false);
        } else {
            if (transpose)
                TA_RESOLVE(uint64_t, true);
            else
                TA_RESOLVE(uint64_t, false);
        }
#undef TA_RESOLVE
        ctx->launches++;
        TA_CUDA(ctx, cudaGetLastError());
        // the filter launched below (shift-and for needles <= 32, else Myers) runs only if the scan gave up
        only_if = gave_up;
        pigeon = needle_len <= 32;
        fo->segs = (uint32_t)qsubs, fo->gran = qgran;
        fo->alt_flag = gave_up, fo->alt_segs = (uint32_t)subs, fo->alt_gran = TA_SEARCH_SUB;
    }
    if (pigeon) {
        // 256-thread blocks: the 32 KB bank-replicated table is shared by 8 warps (TA_PIGEON_THREADS overrides)
        static const int env_pt = getenv("TA_PIGEON_THREADS") ? atoi(getenv("TA_PIGEON_THREADS")) : 0;
        const int pt = env_pt ? env_pt : 256;
        static const int staged = getenv("TA_PIGEON_STAGED") ? atoi(getenv("TA_PIGEON_STAGED")) : 1;
        if (staged || only_if) {  // coalesced cp.async rows through shared memory (default); TA_PIGEON_STAGED=0 = lane-per-segment loads
            const size_t items = n * (size_t)segs;
            const size_t warps = (items + 31) / 32;
            const unsigned blocks = (unsigned)((warps + (pt / 32) - 1) / (pt / 32));
            const size_t smem = 256 * 32 * sizeof(uint32_t) + (size_t)(pt / 32) * 2 * 32 * PIGEON_ROW;
            if (transpose) {
                TA_CUDA(ctx, cudaFuncSetAttribute(search_pigeon_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                search_pigeon_staged_kernel<true><<<blocks, pt, smem, st>>>(needle_dev, needle_len, hay, hay_off, n, k, pieces, (uint32_t)subs, (uint32_t)segs, idx_out, counter, only_if);
            } else {
                TA_CUDA(ctx, cudaFuncSetAttribute(search_pigeon_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                search_pigeon_staged_kernel<false><<<blocks, pt, smem, st>>>(needle_dev, needle_len, hay, hay_off, n, k, pieces, (uint32_t)subs, (uint32_t)segs, idx_out, counter, only_if);
            }
            ctx->launches++;
            TA_CUDA(ctx, cudaGetLastError());
            return TA_OK;
        }
        const dim3 pgrid((unsigned)((n + pt - 1) / pt), (unsigned)segs);
        const size_t smem = 256 * 32 * sizeof(uint32_t);
        if (transpose) {
            TA_CUDA(ctx, cudaFuncSetAttribute(search_pigeon_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            search_pigeon_kernel<true><<<pgrid, pt, smem, st>>>(needle_dev, needle_len, hay, hay_off, n, k, pieces, (uint32_t)subs, idx_out, counter);
        } else {
            TA_CUDA(ctx, cudaFuncSetAttribute(search_pigeon_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            search_pigeon_kernel<false><<<pgrid, pt, smem, st>>>(needle_dev, needle_len, hay, hay_off, n, k, pieces, (uint32_t)subs, idx_out, counter);
        }
    } else if (needle_len <= 32) {
        if (transpose)
            search_filter_kernel<uint32_t, true><<<grid, 128, 0, st>>>(needle_dev, needle_len, hay, hay_off, n, k, (uint32_t)subs, idx_out, counter, only_if);
        else
            search_filter_kernel<uint32_t, false><<<grid, 128, 0, st>>>(needle_dev, needle_len, hay, hay_off, n, k, (uint32_t)subs, idx_out, counter, only_if);
    } else {
        if (transpose)
            search_filter_kernel<uint64_t, true><<<grid, 128, 0, st>>>(needle_dev, needle_len, hay, hay_off, n, k, (uint32_t)subs, idx_out, counter, only_if);
        else
            search_filter_kernel<uint64_t, false><<<grid, 128, 0, st>>>(needle_dev, needle_len, hay, hay_off, n, k, (uint32_t)subs, idx_out, counter, only_if);
    }
    ctx->launches++;
    TA_CUDA(ctx, cudaGetLastError());
    return TA_OK;
}
