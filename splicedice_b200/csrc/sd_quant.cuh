// sd_quant.cuh -- parameters of the CSR aggregation kernels (shared by sd_quant.cu / sd_ir.cu).
#pragma once
#include "sd_common.cuh"

namespace sd {

struct QuantParams {
    int64_t n_junctions;
    int32_t n_samples;
    const int32_t *counts;
    int64_t ld_counts;
    const int32_t *row_ptr;
    const int32_t *col_idx;
    const uint8_t *low_mask;
    int64_t ld_mask;
    float *ps32;
    int64_t ld_ps32;
    double *ps64;
    int64_t ld_ps64;
    int64_t *exc;
    int64_t ld_exc;
    const double *median;    // K4 (ir_table): IR = median / (median + inc + exc); null for PS
    int64_t ld_median;
    double *ir;
    int64_t ld_ir;
    int64_t row_begin, row_end;
    int32_t rows_per_tile;   // R
    int32_t lpr_log2;        // log2(lanes per row); C = 4 << lpr_log2 columns per slab
    int32_t n_slabs;
    int32_t vec_stores;      // outputs are 16-byte aligned with ld % 4 == 0
    int32_t ir_vec;          // median / ir are 16-byte aligned with even ld
};

// Chooses tiled / gather kernel and launches it on `stream` (flags: SD_QUANT_*).
int launch_quant(QuantParams p, uint32_t flags, cudaStream_t stream);

}  // namespace sd
