// sd_peer.cu -- peer-memory plumbing for the multi-GPU pairwise path (one process per GPU).
//
// The per-pair Benjamini-Hochberg correction ranks whole columns of the p-value matrix, so
// row-sharded Fisher results have to reach the column owners (pairwise_fisher.py:186-191 across
// GPUs).  Instead of a separate all-to-all the Fisher kernel stores into the owners' matrices
// directly (sd_fisher_pairwise_scatter); that needs buffers every rank can address.  These entry
// points allocate such a buffer, export it as a CUDA IPC handle (64 opaque bytes the ranks exchange
// through torch.distributed), and map a peer's handle into the calling process: the mapping is a
// device pointer valid on the caller's current device, backed by NVLink peer access.  The adjusted
// column blocks travel back with sd_peer_scatter_rows (one kernel of peer stores).
#include <string.h>

#include "sd_common.cuh"

namespace sd {
namespace {

constexpr int kMaxDest = 16;
struct ScatterRows {
    const double *src;          // [n_rows, width], dense
    int64_t n_rows, width;
    int32_t n_dest;
    int64_t row_begin[kMaxDest + 1];
    double *dest[kMaxDest];     // rank g's [row_begin[g + 1] - row_begin[g], ...] matrix (possibly peer memory)
    int64_t dest_ld[kMaxDest];
    int64_t dest_col0;          // column of dest at which `width` columns of src go
};

// One CTA per 16 source rows; a warp stores 256 contiguous bytes at a time, which is what NVLink
// wants (a strided cudaMemcpy2D of 2 KB rows moved the same bytes four times slower).
__global__ void __launch_bounds__(256) scatter_rows_kernel(const ScatterRows q)
{
    const int64_t r_first = (int64_t)blockIdx.x * 16;
    int g = 0;
    for (int64_t r = r_first; r < min(r_first + 16, q.n_rows); ++r) {
        while (g + 1 < q.n_dest && r >= q.row_begin[g + 1]) ++g;
        const double *src = q.src + r * q.width;
        double *dst = q.dest[g] + (r - q.row_begin[g]) * q.dest_ld[g] + q.dest_col0;
        for (int64_t c = threadIdx.x; c < q.width; c += blockDim.x) dst[c] = __ldcs(src + c);
    }
}

}  // namespace
}  // namespace sd

extern "C" {

int sd_peer_scatter_rows(const double *src, int64_t n_rows, int64_t width, int32_t n_dest, double *const *dest,
                         const int64_t *row_begin, const int64_t *dest_ld, int64_t dest_col0, void *stream)
{
    SD_REQUIRE(n_rows >= 0 && width >= 0, "sd_peer_scatter_rows: negative size");
    if (n_rows == 0 || width == 0) return SD_OK;
    SD_REQUIRE(src && dest && row_begin && dest_ld, "sd_peer_scatter_rows: null pointer");
    SD_REQUIRE(n_dest >= 1 && n_dest <= sd::kMaxDest, "sd_peer_scatter_rows: 1..%d destinations", sd::kMaxDest);
    SD_REQUIRE(row_begin[0] == 0 && row_begin[n_dest] == n_rows, "sd_peer_scatter_rows: the row blocks must cover [0, n_rows)");
    sd::ScatterRows q{};
    q.src = src; q.n_rows = n_rows; q.width = width; q.n_dest = n_dest; q.dest_col0 = dest_col0;
    for (int g = 0; g <= n_dest; ++g) q.row_begin[g] = row_begin[g];
    for (int g = 0; g < n_dest; ++g) {
        SD_REQUIRE(row_begin[g + 1] == row_begin[g] || (dest[g] && dest_ld[g] >= dest_col0 + width),
                   "sd_peer_scatter_rows: destination %d: null pointer or ld too small", g);
        q.dest[g] = dest[g]; q.dest_ld[g] = dest_ld[g];
    }
    const int64_t blocks = (n_rows + 15) / 16;
    SD_REQUIRE(blocks <= 0x7FFFFFFF, "sd_peer_scatter_rows: too many rows");
    sd::scatter_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(q);
    return sd::check_launch("scatter_rows_kernel");
}

int sd_peer_alloc(size_t bytes, void **ptr, unsigned char *handle64)
{
    SD_REQUIRE(ptr && handle64, "sd_peer_alloc: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    void *p = nullptr;
    SD_CHECK_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return sd::fail(SD_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return SD_OK;
}

int sd_peer_free(void *ptr)
{
    if (ptr) SD_CHECK_CUDA(cudaFree(ptr));
    return SD_OK;
}

int sd_peer_open(const unsigned char *handle64, void **ptr)
{
    SD_REQUIRE(ptr && handle64, "sd_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SD_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SD_OK;
}

int sd_peer_close(void *ptr)
{
    if (ptr) SD_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
    return SD_OK;
}

}  // extern "C"
