// lev_band_common.cuh -- what the general-cost kernels share (lev_band.cu: lane group per pair, u32 cells;
// lev_diag16.cu: thread per pair, packed u16 cells): the launch arguments and the per-pair band.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ta_common.cuh"

struct BandArgs {
    const uint8_t *a;
    const uint64_t *a_off;
    const uint8_t *b;
    const uint64_t *b_off;
    const uint32_t *idx;  // optional indirection: work item w -> pair idx[w]
    size_t pair_base;     // without idx: work item w -> pair pair_base + w
    size_t n;
    uint32_t k;
    uint32_t mism, gap, sgap, tcost;
    uint32_t slot;  // shared-memory bytes reserved per string (SMEM variants)
    uint32_t *out;
    // traceback (TRACE variants): one byte per band cell, work item w owns trace[w * trace_stride ..), laid out
    // [anti-diagonal s - s0][cell]; 0 = substitution/match, 1 = a-gap, 2 = b-gap, 3 = transposition
    uint8_t *trace;
    size_t trace_stride;
};

#include "lev_band_info.cuh"

// thread-per-pair u16 kernel (lev_diag16.cu)
bool ta_diag16_can_handle(size_t n, uint32_t k, ta_costs c, uint32_t max_len, uint32_t W);
int ta_launch_lev_diag16(ta_ctx *ctx, const BandArgs &args, ta_costs costs, uint32_t W, cudaStream_t st);
