// sd_peer.cu -- peer-memory plumbing for the multi-GPU pairwise path (one process per GPU).
//
// The per-pair Benjamini-Hochberg correction ranks whole columns of the p-value matrix, so
// row-sharded Fisher results have to reach the column owners (pairwise_fisher.py:186-191 across
// GPUs).  Instead of a separate all-to-all the Fisher kernel stores into the owners' matrices
// directly (sd_fisher_pairwise_scatter); that needs buffers every rank can address.  These entry
// points allocate such a buffer, export it as a CUDA IPC handle (64 opaque bytes the ranks exchange
// through torch.distributed), and map a peer's handle into the calling process: the mapping is a
// device pointer valid on the caller's current device, backed by NVLink peer access.  The adjusted
// column blocks travel back with sd_peer_copy2d (a strided copy on the caller's stream).
#include <string.h>

#include "sd_common.cuh"

extern "C" {

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

int sd_peer_copy2d(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width_bytes, size_t height,
                   void *stream)
{
    if (width_bytes == 0 || height == 0) return SD_OK;
    SD_REQUIRE(dst && src && dst_pitch >= width_bytes && src_pitch >= width_bytes, "sd_peer_copy2d: bad arguments");
    SD_CHECK_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, height, cudaMemcpyDefault, (cudaStream_t)stream));
    return SD_OK;
}

}  // extern "C"
