// sd_api.cu -- library-level entry points: version, thread-local error text, device info,
// and the two probes (FP64 FMA rate, device copy bandwidth) the bench reports beside the
// roofline numbers.
#include <stdarg.h>

#include "sd_common.cuh"

namespace sd {

std::string &last_error()
{
    static thread_local std::string msg;
    return msg;
}

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

// 8 independent FMA chains per thread; 2 flop per FMA.
__global__ void __launch_bounds__(256) fp64_fma_probe(double *sink, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.9999999, c = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;   // never true; keeps the chains alive
}

__global__ void __launch_bounds__(256) copy_probe(const int4 *__restrict__ src, int4 *__restrict__ dst,
                                                  int64_t n)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

}  // namespace sd

extern "C" {

int sd_version(void) { return SD_ABI_VERSION; }

const char *sd_last_error(void) { return sd::last_error().c_str(); }

int sd_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem)
{
    cudaDeviceProp p;
    SD_CHECK_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return SD_OK;
}

int sd_probe_fp64(double *gflops_out, void *stream_)
{
    SD_REQUIRE(gflops_out != nullptr, "sd_probe_fp64: null output");
    cudaStream_t stream = (cudaStream_t)stream_;
    double *sink = nullptr;
    SD_CHECK_CUDA(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    SD_CHECK_CUDA(cudaEventCreate(&e0));
    SD_CHECK_CUDA(cudaEventCreate(&e1));
    const int blocks = sd::kSMs * 8, iters = 4096;
    sd::fp64_fma_probe<<<blocks, 256, 0, stream>>>(sink, 64, 1.0);   // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        SD_CHECK_CUDA(cudaEventRecord(e0, stream));
        sd::fp64_fma_probe<<<blocks, 256, 0, stream>>>(sink, iters, 1.0);
        SD_CHECK_CUDA(cudaEventRecord(e1, stream));
        SD_CHECK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SD_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flop = 2.0 * 64.0 * iters * 256.0 * blocks;
        double g = flop / (ms * 1e-3) / 1e9;
        if (g > best) best = g;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (int rc = sd::check_launch("fp64_fma_probe")) return rc;
    *gflops_out = best;
    return SD_OK;
}

int sd_probe_copy(int64_t bytes, double *gbs_out, void *stream_)
{
    SD_REQUIRE(gbs_out != nullptr && bytes >= 16, "sd_probe_copy: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    int64_t n = bytes / 16;
    int4 *src = nullptr, *dst = nullptr;
    SD_CHECK_CUDA(cudaMalloc(&src, n * 16));
    SD_CHECK_CUDA(cudaMalloc(&dst, n * 16));
    SD_CHECK_CUDA(cudaMemsetAsync(src, 1, n * 16, stream));
    cudaEvent_t e0, e1;
    SD_CHECK_CUDA(cudaEventCreate(&e0));
    SD_CHECK_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        SD_CHECK_CUDA(cudaEventRecord(e0, stream));
        sd::copy_probe<<<sd::kSMs * 16, 256, 0, stream>>>(src, dst, n);
        SD_CHECK_CUDA(cudaEventRecord(e1, stream));
        SD_CHECK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SD_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double g = 2.0 * n * 16 / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(src);
    cudaFree(dst);
    if (int rc = sd::check_launch("copy_probe")) return rc;
    *gbs_out = best;
    return SD_OK;
}

}  // extern "C"
