// sd_bh.cu -- Benjamini-Hochberg adjustment of the Fisher p-value matrix on the device.
//
// Replaces the multiple-test correction of pairwise_fisher.run_with
// (/root/reference/splicedice/pairwise_fisher.py:182-191): statsmodels
// multipletests(method="fdr_bh")[1], per column ("pairwise", the CLI default) or over the whole
// matrix ("all").  statsmodels' fdrcorrection, restated:
//     sorted ascending p_(1..n);  raw_k = p_(k) / (k / n);  adj_k = min_{m >= k} raw_m;  adj > 1 -> 1
// Both divisions are IEEE binary64 (numpy: ps / (arange(1, n + 1) / float(n))), so __ddiv_rn
// reproduces every bit; ties in p need no care (the running minimum from the right gives every
// member of a tie the value of its last member whatever their order).
//
// Device plan (n = rows * cols values, segments = columns or the whole matrix):
//   0. columns only: tiled transpose of p into a column-major copy pT, so that everything a
//      column needs later sits in one contiguous, L2-sized (rows * 8 B) region -- the random
//      gathers / scatters of steps 3 and 5 then hit L2 and a single page instead of striding
//      through the whole matrix (at 4e8 values the row-major form spent 16 ms there, mostly in
//      TLB and sector misses)
//   1. stable LSD radix sort of (p, linear index) by p          -- cub::DeviceRadixSort, 8 passes
//   2. columns only: stable radix sort of that order by column  -- 2 passes over 32-bit pairs;
//      every column's entries are now contiguous and ascending in p
//   3. raw_k per sorted slot + the minimum of every 4,096-slot chunk
//   4. exclusive running minimum of the chunk minima from the right, per segment
//   5. per chunk: running minimum from the right seeded with (4), clip, scatter by index
//   6. columns only: transpose the adjusted column-major matrix into out
// HBM-bound throughout: ~24 B moved per value per radix pass.
#include <cub/cub.cuh>

#include <algorithm>

#include "sd_common.cuh"

namespace sd {
namespace {

constexpr int kBhThreads = 256;
constexpr int kBhItems = 16;
constexpr int kBhChunk = kBhThreads * kBhItems;

// numpy.minimum: NaN if either side is NaN
struct NanMin {
    __device__ __forceinline__ double operator()(double a, double b) const
    {
        if (a != a) return a;
        if (b != b) return b;
        return a < b ? a : b;
    }
};

__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7FF0000000000000ll); }

struct BhParams {
    int64_t n;                 // values
    int64_t seg_len;           // values per segment
    int64_t chunks_per_seg;
    int64_t n_cols;
    const double *p;
    int64_t ld_p;
    double *out;
    int64_t ld_out;
    const double *sorted_p;    // "all": the sorted keys themselves; columns: NULL (gather through idx)
    bool linear;               // idx addresses p / out directly (the column-major copies); else row * n_cols + col
    const uint32_t *idx;       // index of every sorted slot
    double *raw;               // [n]
    double *chunk_min;         // [segments * chunks_per_seg]
    double *chunk_carry;       // same shape: minimum of the later chunks of the segment
};

__global__ void __launch_bounds__(256) bh_init(int64_t n, int64_t n_cols, const double *p, int64_t ld_p, double *keys,
                                               uint32_t *vals)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = i / n_cols, c = i - r * n_cols;
        keys[i] = p[r * ld_p + c];
        vals[i] = (uint32_t)i;
    }
}

// column of every sorted slot; the index is column * n_rows + row (into the column-major copy)
__global__ void __launch_bounds__(256) bh_column_keys(int64_t n, uint32_t n_rows, const uint32_t *idx, uint32_t *col)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) col[i] = idx[i] / n_rows;
}

// pT[c * n_rows + r] = p[r * ld + c] through 32 x 32 shared-memory tiles; the same pass writes the
// sort's key / value inputs (keys = pT, vals = the index into pT)
__global__ void __launch_bounds__(256) bh_transpose_in(int64_t n_rows, int64_t n_cols, const double *__restrict__ p,
                                                       int64_t ld, double *__restrict__ pT, double *__restrict__ keys,
                                                       uint32_t *__restrict__ vals)
{
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
        if (r < n_rows && c < n_cols) tile[ty + 8 * i][tx] = p[r * ld + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t c = c0 + ty + 8 * i, r = r0 + tx;
        if (r < n_rows && c < n_cols) {
            const int64_t lin = c * n_rows + r;
            const double v = tile[tx][ty + 8 * i];
            pT[lin] = v;
            keys[lin] = v;
            vals[lin] = (uint32_t)lin;
        }
    }
}

// out[r * ld + c] = aT[c * n_rows + r]
__global__ void __launch_bounds__(256) bh_transpose_out(int64_t n_rows, int64_t n_cols, const double *__restrict__ aT,
                                                        double *__restrict__ out, int64_t ld)
{
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t c = c0 + ty + 8 * i, r = r0 + tx;
        if (r < n_rows && c < n_cols) tile[ty + 8 * i][tx] = aT[c * n_rows + r];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty + 8 * i, c = c0 + tx;
        if (r < n_rows && c < n_cols) out[r * ld + c] = tile[tx][ty + 8 * i];
    }
}

// slot range of chunk `blockIdx.x`; item e of thread t is slot hi - 1 - (t * kBhItems + e), so a
// forward scan over (t, e) is the running minimum from the right
struct ChunkRange {
    int64_t seg_begin, lo, hi;
    __device__ __forceinline__ ChunkRange(const BhParams &q)
    {
        const int64_t seg = blockIdx.x / q.chunks_per_seg, c = blockIdx.x - seg * q.chunks_per_seg;
        seg_begin = seg * q.seg_len;
        lo = seg_begin + c * kBhChunk;
        hi = min(lo + (int64_t)kBhChunk, seg_begin + q.seg_len);
    }
};

// The gathers / scatters below touch a random 32-byte sector (and, over gigabytes, a random page)
// per value: latency-bound unless many are in flight, so every thread first issues all
// kBhItems index loads, then all its gathers, and only then computes.
__global__ void __launch_bounds__(kBhThreads) bh_raw(const BhParams q)
{
    using Reduce = cub::BlockReduce<double, kBhThreads>;
    __shared__ typename Reduce::TempStorage tmp;
    const ChunkRange ch(q);
    const double len = (double)q.seg_len;
    const uint32_t n_cols = (uint32_t)q.n_cols;
    const double *__restrict__ src = q.p;
    const uint32_t *__restrict__ idx = q.idx;
    double pv[kBhItems];
    // coalesced pass over the chunk (the order inside a chunk does not matter for its minimum)
    if (q.sorted_p) {
#pragma unroll
        for (int e = 0; e < kBhItems; ++e) {
            const int64_t s = ch.lo + threadIdx.x + e * kBhThreads;
            pv[e] = s < ch.hi ? __ldg(q.sorted_p + s) : 0.0;
        }
    } else {
        uint32_t id[kBhItems];
#pragma unroll
        for (int e = 0; e < kBhItems; ++e) {
            const int64_t s = ch.lo + threadIdx.x + e * kBhThreads;
            id[e] = s < ch.hi ? __ldg(idx + s) : 0u;
        }
#pragma unroll
        for (int e = 0; e < kBhItems; ++e) {
            const uint32_t r = id[e] / n_cols, c = id[e] - r * n_cols;
            pv[e] = __ldg(src + (q.linear ? (int64_t)id[e] : (int64_t)r * q.ld_p + c));
        }
    }
    double m = pos_inf();
#pragma unroll
    for (int e = 0; e < kBhItems; ++e) {
        const int64_t s = ch.lo + threadIdx.x + e * kBhThreads;
        if (s < ch.hi) {
            const double ecdf = __ddiv_rn((double)(s - ch.seg_begin + 1), len);
            const double raw = __ddiv_rn(pv[e], ecdf);
            q.raw[s] = raw;
            m = NanMin()(m, raw);
        }
    }
    m = Reduce(tmp).Reduce(m, NanMin());
    if (threadIdx.x == 0) q.chunk_min[blockIdx.x] = m;
}

// one CTA per segment: chunk_carry[c] = min over chunks c' > c of the segment (+inf for the last)
__global__ void __launch_bounds__(kBhThreads) bh_carry(const BhParams q)
{
    using Scan = cub::BlockScan<double, kBhThreads>;
    __shared__ typename Scan::TempStorage tmp;
    const int64_t base = (int64_t)blockIdx.x * q.chunks_per_seg;
    double running = pos_inf();
    for (int64_t done = 0; done < q.chunks_per_seg; done += kBhThreads) {
        const int64_t k = done + threadIdx.x;                        // k-th chunk from the right
        const bool live = k < q.chunks_per_seg;
        const double v = live ? q.chunk_min[base + q.chunks_per_seg - 1 - k] : pos_inf();
        double excl, total;
        Scan(tmp).ExclusiveScan(v, excl, pos_inf(), NanMin(), total);
        if (live) q.chunk_carry[base + q.chunks_per_seg - 1 - k] = NanMin()(running, excl);
        running = NanMin()(running, total);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kBhThreads) bh_finish(const BhParams q)
{
    using Scan = cub::BlockScan<double, kBhThreads>;
    __shared__ typename Scan::TempStorage tmp;
    const ChunkRange ch(q);
    const int64_t first = ch.hi - 1 - (int64_t)threadIdx.x * kBhItems;   // this thread's right-most slot
    const double *__restrict__ raw = q.raw;
    const uint32_t *__restrict__ idx = q.idx;
    const uint32_t n_cols = (uint32_t)q.n_cols;
    double v[kBhItems];
    uint32_t id[kBhItems];
#pragma unroll
    for (int e = 0; e < kBhItems; ++e) {
        const int64_t s = first - e;
        v[e] = s >= ch.lo ? __ldg(raw + s) : pos_inf();
        id[e] = s >= ch.lo ? __ldg(idx + s) : 0u;
    }
    double run = pos_inf();
#pragma unroll
    for (int e = 0; e < kBhItems; ++e) {
        run = NanMin()(run, v[e]);
        v[e] = run;
    }
    double before;
    Scan(tmp).ExclusiveScan(run, before, pos_inf(), NanMin());
    before = NanMin()(before, q.chunk_carry[blockIdx.x]);
#pragma unroll
    for (int e = 0; e < kBhItems; ++e) {
        if (first - e >= ch.lo) {
            double a = NanMin()(before, v[e]);
            if (a > 1.0) a = 1.0;
            const uint32_t r = id[e] / n_cols, c = id[e] - r * n_cols;
            q.out[q.linear ? (int64_t)id[e] : (int64_t)r * q.ld_out + c] = a;
        }
    }
}

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

inline int bits_for(int64_t n)
{
    int b = 1;
    while ((int64_t(1) << b) < n) ++b;
    return b;
}

struct BhWs {
    size_t key_a, key_b, val_a, val_b, p_t, chunk_min, chunk_carry, cub, total;
    size_t cub_bytes;
    int64_t segments, seg_len, chunks_per_seg;
};

int layout(int64_t n_rows, int64_t n_cols, int mode, BhWs *w)
{
    const int64_t n = n_rows * n_cols;
    w->segments = mode == SD_BH_COLUMNS ? n_cols : 1;
    w->seg_len = mode == SD_BH_COLUMNS ? n_rows : n;
    w->chunks_per_seg = (w->seg_len + kBhChunk - 1) / kBhChunk;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    w->key_a = take((size_t)n * 8);
    w->key_b = take((size_t)n * 8);
    w->val_a = take((size_t)n * 4);
    w->val_b = take((size_t)n * 4);
    w->p_t = take(mode == SD_BH_COLUMNS && n_cols > 1 ? (size_t)n * 8 : 0);
    w->chunk_min = take((size_t)(w->segments * w->chunks_per_seg) * 8);
    w->chunk_carry = take((size_t)(w->segments * w->chunks_per_seg) * 8);
    size_t b1 = 0, b2 = 0;
    cub::DoubleBuffer<double> k64(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> k32(nullptr, nullptr), v32(nullptr, nullptr);
    SD_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, k64, v32, (int)n));
    SD_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b2, k32, v32, (int)n));
    w->cub_bytes = std::max(b1, b2);
    w->cub = take(w->cub_bytes);
    w->total = off;
    return SD_OK;
}

inline int grid_for(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, (int64_t)kSMs * 16); }

}  // namespace
}  // namespace sd

extern "C" {

size_t sd_bh_workspace_bytes(int64_t n_rows, int64_t n_cols, int mode)
{
    if (n_rows <= 0 || n_cols <= 0 || n_rows > INT32_MAX / n_cols || (mode != SD_BH_COLUMNS && mode != SD_BH_ALL)) {
        sd::fail(SD_ERR_INVALID, "sd_bh_workspace_bytes: bad shape or mode");
        return 0;
    }
    sd::BhWs w;
    if (sd::layout(n_rows, n_cols, mode, &w) != SD_OK) return 0;
    return w.total;
}

int sd_bh_adjust(int64_t n_rows, int64_t n_cols, const double *p, int64_t ld_p, double *out, int64_t ld_out, int mode,
                 void *workspace, size_t workspace_bytes, void *stream_)
{
    SD_REQUIRE(n_rows >= 0 && n_cols >= 0, "sd_bh_adjust: negative size");
    SD_REQUIRE(mode == SD_BH_COLUMNS || mode == SD_BH_ALL, "sd_bh_adjust: mode must be SD_BH_COLUMNS or SD_BH_ALL");
    if (n_rows == 0 || n_cols == 0) return SD_OK;
    if (n_rows > INT32_MAX / n_cols)
        return sd::fail(SD_ERR_UNSUPPORTED, "sd_bh_adjust: more than 2^31 - 1 values per call (split the columns)");
    SD_REQUIRE(p && out && workspace, "sd_bh_adjust: null pointer");
    SD_REQUIRE(ld_p >= n_cols && ld_out >= n_cols, "sd_bh_adjust: ld too small");
    cudaStream_t stream = (cudaStream_t)stream_;
    sd::BhWs w;
    if (int rc = sd::layout(n_rows, n_cols, mode, &w)) return rc;
    if (workspace_bytes < w.total)
        return sd::fail(SD_ERR_WORKSPACE, "sd_bh_adjust: workspace %zu < %zu bytes", workspace_bytes, w.total);
    SD_REQUIRE(sd::aligned16(workspace), "sd_bh_adjust: workspace must be 16-byte aligned");
    char *base = static_cast<char *>(workspace);
    const int64_t n = n_rows * n_cols;
    double *key_a = reinterpret_cast<double *>(base + w.key_a), *key_b = reinterpret_cast<double *>(base + w.key_b);
    uint32_t *val_a = reinterpret_cast<uint32_t *>(base + w.val_a), *val_b = reinterpret_cast<uint32_t *>(base + w.val_b);
    void *cub_ws = base + w.cub;
    size_t cub_b = w.cub_bytes;

    const bool by_column = mode == SD_BH_COLUMNS && n_cols > 1;
    double *p_t = reinterpret_cast<double *>(base + w.p_t);
    const dim3 tiles((unsigned)((n_rows + 31) / 32), (unsigned)((n_cols + 31) / 32));
    if (by_column) {
        SD_REQUIRE(tiles.y <= 65535u, "sd_bh_adjust: more than 2,097,120 columns per call (split the columns)");
        sd::bh_transpose_in<<<tiles, 256, 0, stream>>>(n_rows, n_cols, p, ld_p, p_t, key_a, val_a);
        if (int rc = sd::check_launch("bh_transpose_in")) return rc;
    } else {
        sd::bh_init<<<sd::grid_for(n), 256, 0, stream>>>(n, n_cols, p, ld_p, key_a, val_a);
        if (int rc = sd::check_launch("bh_init")) return rc;
    }
    cub::DoubleBuffer<double> keys(key_a, key_b);
    cub::DoubleBuffer<uint32_t> vals(val_a, val_b);
    SD_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_b, keys, vals, (int)n, 0, 64, stream));

    sd::BhParams q{};
    q.n = n; q.seg_len = w.seg_len; q.chunks_per_seg = w.chunks_per_seg; q.n_cols = n_cols;
    q.p = p; q.ld_p = ld_p; q.out = out; q.ld_out = ld_out;
    q.chunk_min = reinterpret_cast<double *>(base + w.chunk_min);
    q.chunk_carry = reinterpret_cast<double *>(base + w.chunk_carry);
    if (by_column) {
        // the sorted keys are not needed again (p is gathered from the column-major copy through
        // the index): their two buffers become the column keys and, afterwards, the raw values
        uint32_t *col_a = reinterpret_cast<uint32_t *>(key_a), *col_b = col_a + n;
        sd::bh_column_keys<<<sd::grid_for(n), 256, 0, stream>>>(n, (uint32_t)n_rows, vals.Current(), col_a);
        if (int rc = sd::check_launch("bh_column_keys")) return rc;
        cub::DoubleBuffer<uint32_t> cols(col_a, col_b);
        cub_b = w.cub_bytes;
        SD_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_b, cols, vals, (int)n, 0, sd::bits_for(n_cols), stream));
        q.sorted_p = nullptr;
        q.raw = key_b;
        q.linear = true;                  // gather from / scatter into the column-major copy
        q.p = p_t; q.out = p_t;
    } else {
        q.sorted_p = keys.Current();
        q.raw = keys.Alternate();
        q.linear = false;
    }
    q.idx = vals.Current();
    const int n_chunks = (int)(w.segments * w.chunks_per_seg);
    sd::bh_raw<<<n_chunks, sd::kBhThreads, 0, stream>>>(q);
    if (int rc = sd::check_launch("bh_raw")) return rc;
    sd::bh_carry<<<(int)w.segments, sd::kBhThreads, 0, stream>>>(q);
    if (int rc = sd::check_launch("bh_carry")) return rc;
    sd::bh_finish<<<n_chunks, sd::kBhThreads, 0, stream>>>(q);
    if (int rc = sd::check_launch("bh_finish")) return rc;
    if (by_column) {
        sd::bh_transpose_out<<<tiles, 256, 0, stream>>>(n_rows, n_cols, p_t, out, ld_out);
        return sd::check_launch("bh_transpose_out");
    }
    return SD_OK;
}

}  // extern "C"
