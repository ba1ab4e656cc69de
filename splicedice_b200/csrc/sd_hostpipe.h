// sd_hostpipe.h -- per-device context of the host-buffer entry points (see sd_hostpipe.cu).
#pragma once
#include <cuda_runtime.h>

#include <mutex>

namespace sd {

// Streams and the library's private device-memory pool of one device, locked for one host-buffer call.
struct HostLease {
    std::unique_lock<std::mutex> lock;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaMemPool_t pool = nullptr;
};
int host_lease(int device, HostLease *lease);

}  // namespace sd
