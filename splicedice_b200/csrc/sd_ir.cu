// sd_ir.cu -- K4: intron-retention ratio and the 5-point relative standard deviation.
//
// Replaces the arithmetic of ir_table.calculateIR (/root/reference/splicedice/ir_table.py:118-132).
// The ratio shares the CSR aggregation kernels of sd_quant.cu (same tile staging, a different
// epilogue):  IR = median / (median + inc + sum of inc over the adjacency list).
#include "sd_common.cuh"
#include "sd_quant.cuh"

namespace sd {

// numpy: np.std(cov) / np.mean(cov) over 5 points (ir_table.py:118-120).  mean = (sequential
// sum) / 5; std = sqrt(sum((x - mean)^2) / 5); every operation rounded separately, as numpy
// does (no fused multiply-add).
__global__ void __launch_bounds__(256) rsd5_kernel(int64_t n, const double *__restrict__ cov, double *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double *c = cov + i * 5;
        double x[5];
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 5; ++k) { x[k] = c[k]; s = __dadd_rn(s, x[k]); }
        const double mean = __ddiv_rn(s, 5.0);
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double dlt = __dsub_rn(x[k], mean);
            v = __dadd_rn(v, __dmul_rn(dlt, dlt));
        }
        const double sd_ = __dsqrt_rn(__ddiv_rn(v, 5.0));
        out[i] = (mean == 0.0 && sd_ == 0.0) ? __longlong_as_double((long long)0xFFF8000000000000ull)
                                             : __ddiv_rn(sd_, mean);
    }
}

}  // namespace sd

extern "C" {

int sd_ir_ratio(int64_t n_junctions, int32_t n_samples, const double *median, int64_t ld_median,
                const int32_t *counts, int64_t ld_counts, const int32_t *row_ptr, const int32_t *col_idx,
                double *ir_out, int64_t ld_ir, int64_t row_begin, int64_t row_end, void *stream)
{
    SD_REQUIRE(n_junctions >= 0 && n_samples >= 0, "sd_ir_ratio: negative size");
    SD_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= n_junctions, "sd_ir_ratio: bad row range");
    if (row_begin == row_end || n_samples == 0) return SD_OK;
    SD_REQUIRE(median && counts && ir_out, "sd_ir_ratio: null pointer");
    SD_REQUIRE(ld_median >= n_samples && ld_counts >= n_samples && ld_ir >= n_samples, "sd_ir_ratio: ld too small");
    sd::QuantParams p{};
    p.n_junctions = n_junctions; p.n_samples = n_samples;
    p.counts = counts; p.ld_counts = ld_counts;
    p.row_ptr = row_ptr; p.col_idx = col_idx;
    p.median = median; p.ld_median = ld_median;
    p.ir = ir_out; p.ld_ir = ld_ir;
    p.row_begin = row_begin; p.row_end = row_end;
    return sd::launch_quant(p, SD_QUANT_AUTO, (cudaStream_t)stream);
}

int sd_rsd5(int64_t n, const double *cov5, double *rsd_out, void *stream)
{
    SD_REQUIRE(n >= 0, "sd_rsd5: negative size");
    if (n == 0) return SD_OK;
    SD_REQUIRE(cov5 && rsd_out, "sd_rsd5: null pointer");
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sd::kSMs * 32);
    sd::rsd5_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, cov5, rsd_out);
    return sd::check_launch("rsd5_kernel");
}

}  // extern "C"
