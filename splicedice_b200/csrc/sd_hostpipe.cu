// sd_hostpipe.cu -- sd_quant_ps_host: the PS path for HOST buffers (the call a ctypes user of the
// reference-facing API makes: host counts in, host PS out, SPLICEDICE.py:297-310).
//
// The matrix crosses PCIe once each way, so the call is bound by the link, not by the kernel
// (1.6 GB each way at 400,000 x 1,000 against a 0.5 ms kernel).  What this file does about it:
//   * row blocks flow through three streams (H2D, kernel, D2H) kept on a per-device context that
//     lives across calls together with its events and a PRIVATE memory pool for the device
//     buffers (the process-wide default pool is left alone);
//   * a row block's kernel is launched as soon as the furthest block its adjacency reaches has
//     landed, its PS block is copied back as soon as the kernel is done;
//   * counts below 65,536 -- every RNA-seq junction count in practice -- cross the link as uint16:
//     host threads narrow block b + 1 into a ring of pinned staging slots while block b is in
//     flight, and a small kernel widens it back to int32 in HBM.  That halves the H2D bytes; on a
//     full-duplex link that is already at its combined ceiling the D2H direction then runs at its
//     own rate.  The decision is per block (one value >= 65,536 or < 0 sends that block as int32),
//     so results never depend on it.
#include <immintrin.h>
#include <sched.h>

#include <chrono>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "sd_common.cuh"
#include "sd_hostpipe.h"
#include "sd_quant.cuh"

namespace sd {
namespace {

// ---- host worker threads (process-wide, started on first use) ------------------------------
class Workers {
public:
    static Workers &get()
    {
        static Workers w;
        return w;
    }
    int size() const { return (int)threads_.size(); }
    void submit(std::function<void()> fn)
    {
        {
            std::lock_guard<std::mutex> lock(mu_);
            queue_.push_back(std::move(fn));
        }
        cv_.notify_one();
    }

private:
    Workers()
    {
        int n = 0;
        if (const char *env = getenv("SD_HOST_THREADS")) n = atoi(env);
        if (n <= 0) {
            cpu_set_t set;
            int cpus = (int)std::thread::hardware_concurrency();
            if (sched_getaffinity(0, sizeof set, &set) == 0) cpus = CPU_COUNT(&set);
            // half of the CPUs this process may use, shared between the ranks of one torchrun job; more
            // than 8 threads only adds host-memory traffic that slows the D2H stream (measured on a
            // 16-vCPU B200 box, 400,000 x 1,000: 4 / 8 / 12 / 16 threads = 41.2 / 32.8 / 33.9 / 34.3 ms)
            int ranks = 1;
            if (const char *lw = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(lw));
            n = std::min(8, std::max(2, cpus / 2 / ranks));
        }
        for (int i = 0; i < n; ++i) threads_.emplace_back([this] { loop(); });
    }
    ~Workers()
    {
        {
            std::lock_guard<std::mutex> lock(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    void loop()
    {
        for (;;) {
            std::function<void()> fn;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_.wait(lock, [this] { return stop_ || !queue_.empty(); });
                if (queue_.empty()) return;
                fn = std::move(queue_.front());
                queue_.pop_front();
            }
            fn();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> queue_;
    std::vector<std::thread> threads_;
    bool stop_ = false;
};

// rows [r0, r1) of an int32 matrix -> uint16 rows of `ld16` elements (padding zeroed); returns the
// OR of every value seen (any bit at or above 2^16, sign bit included, means "does not fit").
// When source and destination rows are both dense (ld_src == ld16 == n) the chunk is one flat
// array: 16 values per step, packed with VPACKUSDW and written with non-temporal stores (the
// staging slot is read next by the DMA engine, never by this core; the chunk starts 32-byte
// aligned because chunks begin on multiples of 16 rows).
__attribute__((target("avx2"))) uint32_t narrow_rows_avx2(const int32_t *src, int64_t ld_src, uint16_t *dst,
                                                            int64_t ld16, int64_t r0, int64_t r1, int32_t n)
{
    uint32_t seen = 0;
    if (ld_src == n && ld16 == n && ((uintptr_t)dst & 31u) == 0) {
        const int32_t *s = src + r0 * ld_src;
        const int64_t total = (r1 - r0) * (int64_t)n;
        __m256i acc = _mm256_setzero_si256();
        int64_t i = 0;
        for (; i + 32 <= total; i += 32) {
            const __m256i a = _mm256_loadu_si256((const __m256i *)(s + i));
            const __m256i b = _mm256_loadu_si256((const __m256i *)(s + i + 8));
            const __m256i c = _mm256_loadu_si256((const __m256i *)(s + i + 16));
            const __m256i d = _mm256_loadu_si256((const __m256i *)(s + i + 24));
            acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(c, d)));
            const __m256i lo = _mm256_permute4x64_epi64(_mm256_packus_epi32(a, b), 0xD8);
            const __m256i hi = _mm256_permute4x64_epi64(_mm256_packus_epi32(c, d), 0xD8);
            _mm256_stream_si256((__m256i *)(dst + i), lo);
            _mm256_stream_si256((__m256i *)(dst + i + 16), hi);
        }
        alignas(32) uint32_t lanes[8];
        _mm256_store_si256((__m256i *)lanes, acc);
        for (int k = 0; k < 8; ++k) seen |= lanes[k];
        for (; i < total; ++i) {
            seen |= (uint32_t)s[i];
            dst[i] = (uint16_t)s[i];
        }
        _mm_sfence();
        return seen;
    }
    for (int64_t r = r0; r < r1; ++r) {
        const int32_t *s = src + r * ld_src;
        uint16_t *d = dst + (r - r0) * ld16;
        uint32_t acc = 0;
        for (int32_t i = 0; i < n; ++i) {
            acc |= (uint32_t)s[i];
            d[i] = (uint16_t)s[i];
        }
        for (int64_t i = n; i < ld16; ++i) d[i] = 0;
        seen |= acc;
    }
    return seen;
}
uint32_t narrow_rows_plain(const int32_t *src, int64_t ld_src, uint16_t *dst, int64_t ld16, int64_t r0, int64_t r1,
                           int32_t n)
{
    uint32_t seen = 0;
    for (int64_t r = r0; r < r1; ++r) {
        const int32_t *s = src + r * ld_src;
        uint16_t *d = dst + (r - r0) * ld16;
        for (int32_t i = 0; i < n; ++i) {
            seen |= (uint32_t)s[i];
            d[i] = (uint16_t)s[i];
        }
        for (int64_t i = n; i < ld16; ++i) d[i] = 0;
    }
    return seen;
}

// uint16 rows (ld16 elements, ld16 % 4 == 0) -> int32 rows (ldd == ld16 elements)
__global__ void widen_u16_kernel(const uint16_t *__restrict__ src, int32_t *__restrict__ dst, int64_t n_quads)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_quads; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 v = __ldcs(reinterpret_cast<const uint2 *>(src) + i);
        int4 o;
        o.x = (int)(v.x & 0xFFFFu); o.y = (int)(v.x >> 16);
        o.z = (int)(v.y & 0xFFFFu); o.w = (int)(v.y >> 16);
        reinterpret_cast<int4 *>(dst)[i] = o;
    }
}

// ---- per-device context, alive across calls -----------------------------------------------------
constexpr int kMaxDevices = 64;
constexpr int kRingSlots = 4;

struct HostPipe {
    std::mutex mu;                     // one host-buffer call at a time per device
    bool ready = false;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaMemPool_t pool = nullptr;      // private: cached blocks are ours, not the process's
    std::vector<cudaEvent_t> events;
    uint16_t *ring = nullptr;          // pinned staging, kRingSlots slots
    size_t ring_slot_bytes = 0;
    double ms_per_mb[2] = {0, 0};      // best large call as int32 [0] / uint16 [1]: the format autotuner
    int tune_samples[2] = {0, 0};
    long tune_calls = 0;
};
HostPipe g_pipe[kMaxDevices];

int pipe_init(HostPipe &hp, int device)
{
    if (hp.ready) return SD_OK;
    SD_CHECK_CUDA(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
    SD_CHECK_CUDA(cudaStreamCreateWithFlags(&hp.s_k, cudaStreamNonBlocking));
    SD_CHECK_CUDA(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    SD_CHECK_CUDA(cudaMemPoolCreate(&hp.pool, &props));
    uint64_t keep = ~0ull;             // repeated calls reuse the blocks; sd_host_pipeline_trim() releases them
    SD_CHECK_CUDA(cudaMemPoolSetAttribute(hp.pool, cudaMemPoolAttrReleaseThreshold, &keep));
    hp.ready = true;
    return SD_OK;
}

int pipe_events(HostPipe &hp, size_t n)
{
    while (hp.events.size() < n) {
        cudaEvent_t e = nullptr;
        SD_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        hp.events.push_back(e);
    }
    return SD_OK;
}

int pipe_ring(HostPipe &hp, size_t slot_bytes)
{
    if (hp.ring && hp.ring_slot_bytes >= slot_bytes) return SD_OK;
    if (hp.ring) cudaFreeHost(hp.ring);
    hp.ring = nullptr;
    hp.ring_slot_bytes = 0;
    SD_CHECK_CUDA(cudaHostAlloc(&hp.ring, slot_bytes * kRingSlots, cudaHostAllocDefault));
    hp.ring_slot_bytes = slot_bytes;
    return SD_OK;
}

// completion counter of one block's narrowing tasks
struct NarrowJob {
    std::mutex mu;
    std::condition_variable cv;
    int pending = 0;
    uint32_t seen = 0;
    void done(uint32_t s)
    {
        std::lock_guard<std::mutex> lock(mu);
        seen |= s;
        if (--pending == 0) cv.notify_all();
    }
    uint32_t wait()
    {
        std::unique_lock<std::mutex> lock(mu);
        cv.wait(lock, [this] { return pending == 0; });
        return seen;
    }

};

}  // namespace

// The per-device streams and the private pool for the other host-buffer entry point
// (sd_fisher_pairwise_host): holds the device's lock until the lease is destroyed.
int host_lease(int device, HostLease *lease)
{
    if (device < 0 || device >= kMaxDevices) return fail(SD_ERR_INVALID, "device ordinal %d out of range", device);
    HostPipe &hp = g_pipe[device];
    lease->lock = std::unique_lock<std::mutex>(hp.mu);
    if (int rc = pipe_init(hp, device)) return rc;
    lease->s_in = hp.s_in; lease->s_k = hp.s_k; lease->s_out = hp.s_out; lease->pool = hp.pool;
    return SD_OK;
}

}  // namespace sd

extern "C" {

int sd_host_pipeline_trim(int device)
{
    using namespace sd;
    SD_REQUIRE(device >= 0 && device < kMaxDevices, "sd_host_pipeline_trim: device ordinal %d out of range", device);
    HostPipe &hp = g_pipe[device];
    std::lock_guard<std::mutex> lock(hp.mu);
    if (!hp.ready) return SD_OK;
    int prev = 0;
    SD_CHECK_CUDA(cudaGetDevice(&prev));
    SD_CHECK_CUDA(cudaSetDevice(device));
    cudaStreamSynchronize(hp.s_in);
    cudaStreamSynchronize(hp.s_k);
    cudaStreamSynchronize(hp.s_out);
    cudaMemPoolTrimTo(hp.pool, 0);
    if (hp.ring) cudaFreeHost(hp.ring);
    hp.ring = nullptr;
    hp.ring_slot_bytes = 0;
    cudaSetDevice(prev);
    return SD_OK;
}

int sd_quant_ps_host(int device, int64_t n_junctions, int32_t n_samples, const int32_t *counts,
                     int64_t ld_counts, const int32_t *row_ptr, const int32_t *col_idx,
                     const uint8_t *low_mask, int64_t ld_mask, float *ps_f32, int64_t ld_ps32)
{
    using namespace sd;
    SD_REQUIRE(n_junctions >= 0 && n_samples >= 0, "sd_quant_ps_host: negative size");
    if (n_junctions == 0 || n_samples == 0) return SD_OK;
    SD_REQUIRE(counts && row_ptr && ps_f32, "sd_quant_ps_host: null pointer");
    SD_REQUIRE(ld_counts >= n_samples && ld_ps32 >= n_samples, "sd_quant_ps_host: ld < n_samples");
    SD_REQUIRE(!low_mask || ld_mask >= n_samples, "sd_quant_ps_host: ld_mask < n_samples");
    SD_REQUIRE(device >= 0 && device < kMaxDevices, "sd_quant_ps_host: device ordinal %d out of range", device);
    const int64_t J = n_junctions;
    const int64_t nnz = row_ptr[J];
    SD_REQUIRE(nnz >= 0 && (nnz == 0 || col_idx), "sd_quant_ps_host: bad CSR");

    const int64_t ldd = (n_samples + 3) & ~(int64_t)3;           // device leading dimension
    const int64_t ldm = (n_samples + 15) & ~(int64_t)15;
    // The link format.  uint16 staging halves the H2D bytes at the price of one pass of host threads
    // over the matrix (1.6 GB read + 0.8 GB written at 400,000 x 1,000).  Whether that pays depends on
    // the host: on one 16-vCPU B200 box it took 32.7 ms against 36.4 ms for plain int32 (duplex
    // ceiling of the same bytes 33.1 ms, D2H alone 28.2 ms); on two other boxes of the same pool the
    // host threads slowed the copy-back stream and it took 40-42 ms against 35.7; with two ranks
    // sharing a host it loses outright (60 against 48.9 ms, which IS the two-GPU duplex ceiling).  So
    // the format is chosen by measurement: ranks of a multi-GPU job (LOCAL_WORLD_SIZE > 1) and small
    // matrices send int32; otherwise the per-device context times its first two large calls in each
    // format and keeps the faster one, trying the other again every 32nd call.  SD_QUANT_HOST_U16=0/1
    // overrides.  Results never depend on the format.
    HostPipe &hp = g_pipe[device];
    std::lock_guard<std::mutex> pipe_lock(hp.mu);
    const double cells_mb = (double)J * n_samples * 4.0 / 1048576.0;
    bool use_u16 = false, tuning = false;
    if (const char *env = getenv("SD_QUANT_HOST_U16")) {
        use_u16 = atoi(env) != 0;
    } else {
        bool alone = true;
        if (const char *lw = getenv("LOCAL_WORLD_SIZE")) alone = atoi(lw) <= 1;
        if (alone && cells_mb >= 256.0) {
            tuning = true;
            ++hp.tune_calls;
            // two samples per format before deciding (a format's first call also pays first-use costs:
            // module load, tensor-map entry point, the pinned ring), the better of the two counts
            if (hp.tune_samples[0] < 2) use_u16 = false;
            else if (hp.tune_samples[1] < 2) use_u16 = true;
            else {
                use_u16 = hp.ms_per_mb[1] < hp.ms_per_mb[0];
                if (hp.tune_calls % 32 == 0) use_u16 = !use_u16;          // look at the other format again
            }
        }
    }
    // row blocks of ~16 MB of int32 (8 MB on the link as uint16; 32 MB when everything crosses as
    // int32): 8 / 16 / 32 MB measured 33.4 / 32.8 / 33.3 ms with uint16, 38.6 / 36.4 / 36.3 ms
    // without.  Tuning knob for experiments: SD_QUANT_HOST_BLOCK_MB
    int64_t block_mb = use_u16 ? 16 : 32;
    if (const char *env = getenv("SD_QUANT_HOST_BLOCK_MB")) block_mb = std::max<int64_t>(1, atoll(env));
    int64_t block_rows = std::max<int64_t>(64, (block_mb << 20) / (ldd * 4));
    block_rows = (block_rows + 63) & ~(int64_t)63;
    const int64_t n_blocks = (J + block_rows - 1) / block_rows;

    // furthest block each row block's adjacency reaches
    std::vector<int64_t> need(n_blocks);
    for (int64_t b = 0; b < n_blocks; ++b) {
        int64_t r0 = b * block_rows, r1 = std::min(J, r0 + block_rows);
        int32_t hi = (int32_t)(r1 - 1);
        for (int64_t k = row_ptr[r0]; k < row_ptr[r1]; ++k) {
            int32_t c = col_idx[k];
            if (c < 0 || c >= J)
                return fail(SD_ERR_INVALID, "sd_quant_ps_host: col_idx[%lld] = %d out of range", (long long)k, c);
            hi = std::max(hi, c);
        }
        need[b] = hi / block_rows;
    }

    Workers *workers = use_u16 ? &Workers::get() : nullptr;
    const bool avx2 = __builtin_cpu_supports("avx2");

    int prev_dev = 0;
    SD_CHECK_CUDA(cudaGetDevice(&prev_dev));
    SD_CHECK_CUDA(cudaSetDevice(device));

    int32_t *d_counts = nullptr, *d_row_ptr = nullptr, *d_col = nullptr;
    uint16_t *d_stage = nullptr;
    float *d_ps = nullptr;
    uint8_t *d_mask = nullptr;
    std::vector<NarrowJob> jobs(use_u16 ? n_blocks : 0);
    int64_t submitted = 0;
    int rc = SD_OK;
    // every exit path: wait for the workers still writing the ring, drain the three streams, then
    // hand the device buffers back to the pool
    auto cleanup = [&]() {
        for (int64_t b = 0; b < submitted; ++b) jobs[b].wait();
        if (hp.s_in) cudaStreamSynchronize(hp.s_in);
        if (hp.s_k) cudaStreamSynchronize(hp.s_k);
        if (hp.s_out) cudaStreamSynchronize(hp.s_out);
        if (d_counts) cudaFreeAsync(d_counts, hp.s_out);
        if (d_stage) cudaFreeAsync(d_stage, hp.s_out);
        if (d_ps) cudaFreeAsync(d_ps, hp.s_out);
        if (d_row_ptr) cudaFreeAsync(d_row_ptr, hp.s_out);
        if (d_col) cudaFreeAsync(d_col, hp.s_out);
        if (d_mask) cudaFreeAsync(d_mask, hp.s_out);
        if (hp.s_out) cudaStreamSynchronize(hp.s_out);
        cudaSetDevice(prev_dev);
    };
#define SD_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            rc = sd::fail(SD_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                              \
            cleanup();                                                                            \
            return rc;                                                                            \
        }                                                                                         \
    } while (0)
#define SD_TRY_RC(expr)              \
    do {                             \
        rc = (expr);                 \
        if (rc != SD_OK) {           \
            cleanup();               \
            return rc;               \
        }                            \
    } while (0)

    SD_TRY_RC(pipe_init(hp, device));
    SD_TRY_RC(pipe_events(hp, (size_t)n_blocks * 3));
    cudaEvent_t *ev_copy = hp.events.data(), *ev_in = ev_copy + n_blocks, *ev_k = ev_in + n_blocks;
    const size_t slot_bytes = (size_t)block_rows * ldd * 2;
    if (use_u16) SD_TRY_RC(pipe_ring(hp, slot_bytes));
    cudaStream_t s_in = hp.s_in, s_k = hp.s_k, s_out = hp.s_out;
    SD_TRY(cudaMallocFromPoolAsync(&d_counts, (size_t)J * ldd * 4, hp.pool, s_in));
    SD_TRY(cudaMallocFromPoolAsync(&d_ps, (size_t)J * ldd * 4, hp.pool, s_in));
    SD_TRY(cudaMallocFromPoolAsync(&d_row_ptr, (size_t)(J + 1) * 4, hp.pool, s_in));
    SD_TRY(cudaMallocFromPoolAsync(&d_col, (size_t)std::max<int64_t>(nnz, 1) * 4, hp.pool, s_in));
    if (use_u16) SD_TRY(cudaMallocFromPoolAsync(&d_stage, (size_t)n_blocks * slot_bytes, hp.pool, s_in));
    if (low_mask) SD_TRY(cudaMallocFromPoolAsync(&d_mask, (size_t)J * ldm, hp.pool, s_in));
    // the kernel and copy-back streams must not run ahead of the allocations made on s_in
    SD_TRY(cudaEventRecord(ev_k[0], s_in));
    SD_TRY(cudaStreamWaitEvent(s_k, ev_k[0], 0));
    SD_TRY(cudaStreamWaitEvent(s_out, ev_k[0], 0));
    SD_TRY(cudaMemcpyAsync(d_row_ptr, row_ptr, (size_t)(J + 1) * 4, cudaMemcpyHostToDevice, s_in));
    if (nnz) SD_TRY(cudaMemcpyAsync(d_col, col_idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, s_in));

    auto submit_narrow = [&](int64_t b) {
        const int64_t r0 = b * block_rows, r1 = std::min(J, r0 + block_rows);
        uint16_t *slot = hp.ring + (size_t)(b % kRingSlots) * (hp.ring_slot_bytes / 2);
        const int parts = std::max(1, workers->size());
        const int64_t per = (((r1 - r0 + parts - 1) / parts) + 15) & ~(int64_t)15;   // chunks start 32-byte aligned
        NarrowJob &job = jobs[b];
        job.pending = (int)((r1 - r0 + per - 1) / per);
        for (int64_t a = r0; a < r1; a += per) {
            const int64_t e = std::min(r1, a + per);
            workers->submit([=, &job] {
                uint16_t *dst = slot + (a - r0) * ldd;
                job.done(avx2 ? narrow_rows_avx2(counts, ld_counts, dst, ldd, a, e, n_samples)
                              : narrow_rows_plain(counts, ld_counts, dst, ldd, a, e, n_samples));
            });
        }
    };

    const bool debug = getenv("SD_HOST_PIPE_DEBUG") != nullptr;
    using clk = std::chrono::steady_clock;
    const auto t_begin = clk::now();
    double wait_narrow_ms = 0, wait_slot_ms = 0;
    int64_t wide_blocks = 0;
    int64_t next_kernel = 0;
    for (int64_t b = 0; b < n_blocks; ++b) {
        const int64_t r0 = b * block_rows, r1 = std::min(J, r0 + block_rows);
        bool narrow_ok = false;
        if (use_u16) {
            // keep the narrowing up to kRingSlots - 1 blocks ahead of the copies
            while (submitted < n_blocks && submitted < b + kRingSlots) {
                const auto t0 = clk::now();
                if (submitted >= kRingSlots) SD_TRY(cudaEventSynchronize(ev_copy[submitted - kRingSlots]));   // slot free again
                wait_slot_ms += std::chrono::duration<double, std::milli>(clk::now() - t0).count();
                submit_narrow(submitted);
                ++submitted;
            }
            const auto t0 = clk::now();
            narrow_ok = (jobs[b].wait() >> 16) == 0;
            wait_narrow_ms += std::chrono::duration<double, std::milli>(clk::now() - t0).count();
        }
        if (narrow_ok) {
            const uint16_t *slot = hp.ring + (size_t)(b % kRingSlots) * (hp.ring_slot_bytes / 2);
            uint16_t *d_slot = d_stage + (size_t)b * (slot_bytes / 2);
            SD_TRY(cudaMemcpyAsync(d_slot, slot, (size_t)(r1 - r0) * ldd * 2, cudaMemcpyHostToDevice, s_in));
            SD_TRY(cudaEventRecord(ev_copy[b], s_in));
            const int64_t quads = (r1 - r0) * ldd / 4;
            widen_u16_kernel<<<(unsigned)std::min<int64_t>((quads + 255) / 256, 4 * (int64_t)kSMs), 256, 0, s_in>>>(
                d_slot, d_counts + r0 * ldd, quads);
            SD_TRY_RC(check_launch("widen_u16_kernel"));
        } else {
            if (ld_counts == ldd)
                SD_TRY(cudaMemcpyAsync(d_counts + r0 * ldd, counts + r0 * ld_counts,
                                       (size_t)((r1 - r0 - 1) * ldd + n_samples) * 4, cudaMemcpyHostToDevice, s_in));
            else
                SD_TRY(cudaMemcpy2DAsync(d_counts + r0 * ldd, (size_t)ldd * 4, counts + r0 * ld_counts,
                                         (size_t)ld_counts * 4, (size_t)n_samples * 4, (size_t)(r1 - r0),
                                         cudaMemcpyHostToDevice, s_in));
            if (use_u16) SD_TRY(cudaEventRecord(ev_copy[b], s_in));
            wide_blocks += use_u16 ? 1 : 0;
        }
        if (low_mask)
            SD_TRY(cudaMemcpy2DAsync(d_mask + r0 * ldm, (size_t)ldm, low_mask + r0 * ld_mask, (size_t)ld_mask,
                                     (size_t)n_samples, (size_t)(r1 - r0), cudaMemcpyHostToDevice, s_in));
        SD_TRY(cudaEventRecord(ev_in[b], s_in));

        // every row block whose adjacency is now fully on the device: kernel, then copy back
        while (next_kernel < n_blocks && need[next_kernel] <= b) {
            const int64_t kb = next_kernel++;
            const int64_t k0 = kb * block_rows, k1 = std::min(J, k0 + block_rows);
            SD_TRY(cudaStreamWaitEvent(s_k, ev_in[need[kb]], 0));
            QuantParams p{};
            p.n_junctions = J; p.n_samples = n_samples;
            p.counts = d_counts; p.ld_counts = ldd;
            p.row_ptr = d_row_ptr; p.col_idx = d_col;
            p.low_mask = d_mask; p.ld_mask = ldm;
            p.ps32 = d_ps; p.ld_ps32 = ldd;
            p.row_begin = k0; p.row_end = k1;
            SD_TRY_RC(launch_quant(p, SD_QUANT_TILED, s_k));
            SD_TRY(cudaEventRecord(ev_k[kb], s_k));
            SD_TRY(cudaStreamWaitEvent(s_out, ev_k[kb], 0));
            if (ld_ps32 == ldd)
                SD_TRY(cudaMemcpyAsync(ps_f32 + k0 * ld_ps32, d_ps + k0 * ldd,
                                       (size_t)((k1 - k0 - 1) * ldd + n_samples) * 4, cudaMemcpyDeviceToHost, s_out));
            else
                SD_TRY(cudaMemcpy2DAsync(ps_f32 + k0 * ld_ps32, (size_t)ld_ps32 * 4, d_ps + k0 * ldd, (size_t)ldd * 4,
                                         (size_t)n_samples * 4, (size_t)(k1 - k0), cudaMemcpyDeviceToHost, s_out));
        }
    }
#undef SD_TRY
#undef SD_TRY_RC
    const auto t_enqueued = clk::now();
    cleanup();
    if (tuning) {
        const int m = use_u16 ? 1 : 0;
        const double t = std::chrono::duration<double, std::milli>(clk::now() - t_begin).count() / cells_mb;
        // keep the best of the first samples; later calls track the current speed of the chosen format
        hp.ms_per_mb[m] = hp.tune_samples[m] == 1 ? std::min(hp.ms_per_mb[m], t) : t;
        ++hp.tune_samples[m];
    }
    if (debug)
        fprintf(stderr,
                "[sd_quant_ps_host] %lld blocks of %lld rows, uint16 %s (%lld blocks sent as int32), %d workers: enqueue "
                "loop %.2f ms (waiting for narrowing %.2f, for ring slots %.2f), drain %.2f ms\n",
                (long long)n_blocks, (long long)block_rows, use_u16 ? "on" : "off", (long long)wide_blocks,
                workers ? workers->size() : 0, std::chrono::duration<double, std::milli>(t_enqueued - t_begin).count(),
                wait_narrow_ms, wait_slot_ms, std::chrono::duration<double, std::milli>(clk::now() - t_enqueued).count());
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sd::fail(SD_ERR_CUDA, "sd_quant_ps_host: %s", cudaGetErrorString(e));
    return SD_OK;
}

}  // extern "C"
