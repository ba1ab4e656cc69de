// sd_synth.cu -- device twin of splicedice_b200/synth.py:counts_host (bench / test inputs).
// Integer-only counter-based generator keyed by (seed, row, col): any slab of a matrix far
// larger than host memory can be re-made on the CPU, bit for bit, for parity checks.
#include <math.h>

#include "sd_common.cuh"

namespace sd {

__constant__ int32_t c_log2_lut[256];

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) synth_counts_kernel(uint64_t seed, int64_t row0, int64_t n_rows,
                                                           int32_t n_cols, int64_t logical_cols,
                                                           uint32_t scale, int32_t *out, int64_t ld_out)
{
    const int64_t cells = n_rows * n_cols;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < cells; g += stride) {
        const int64_t r = g / n_cols;
        const int32_t c = (int32_t)(g - r * n_cols);
        const uint64_t ctr = (uint64_t)(row0 + r) * (uint64_t)logical_cols + (uint64_t)c;
        const uint64_t h = mix64(ctr * 0x9E3779B97F4A7C15ull + seed * 0x94D049BB133111EBull + 1ull);
        const uint32_t u = (uint32_t)(h >> 32) | 1u;
        const int lz = __clz((int)u);
        const uint32_t norm = u << lz;
        const uint32_t frac = (norm >> 23) & 0xFFu;
        const int64_t L = ((int64_t)(lz + 1) << 16) - c_log2_lut[frac];
        out[r * ld_out + c] = (int32_t)((L * (int64_t)scale) >> 32);
    }
}

}  // namespace sd

extern "C" int sd_synth_counts(uint64_t seed, int64_t row0, int64_t n_rows, int32_t n_cols,
                               int64_t logical_cols, uint32_t scale, int32_t *out, int64_t ld_out,
                               void *stream)
{
    SD_REQUIRE(n_rows >= 0 && n_cols >= 0 && ld_out >= n_cols, "sd_synth_counts: bad shape");
    if (n_rows == 0 || n_cols == 0) return SD_OK;
    SD_REQUIRE(out != nullptr, "sd_synth_counts: null output");
    int32_t lut[256];
    for (int f = 0; f < 256; ++f) lut[f] = (int32_t)nearbyint(65536.0 * log2(1.0 + f / 256.0));
    SD_CHECK_CUDA(cudaMemcpyToSymbolAsync(sd::c_log2_lut, lut, sizeof lut, 0, cudaMemcpyHostToDevice,
                                          (cudaStream_t)stream));
    const int64_t cells = n_rows * n_cols;
    const int blocks = (int)std::min<int64_t>((cells + 255) / 256, (int64_t)sd::kSMs * 32);
    sd::synth_counts_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(seed, row0, n_rows, n_cols, logical_cols,
                                                                       scale, out, ld_out);
    return sd::check_launch("synth_counts_kernel");
}
