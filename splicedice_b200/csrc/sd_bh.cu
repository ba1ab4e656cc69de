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
// HBM-bound throughout: ~24 B moved per value per radix pass -- 10 passes for the column mode.
//
// Column mode, 8,192 <= rows <= 2^19 (the `pairwise` default on any real matrix): a sample sort
// per column instead (namespace colsort below), ~100 B per value in all:
//   0. the same transpose into column-major pT
//   A. per column: up to 8,192 pseudo-randomly placed samples sorted in shared memory give B - 1
//      splitters (B = 32..1,024 so that a bucket holds ~rows / B <= 512 values); every splitter
//      owns an "equality" bucket, so a value that fills a large part of the column (p = 1 of the
//      zero-margin tables) never makes a bucket that must be sorted
//   B. bucket of every value (binary search in shared memory), counts per (column, bucket)
//   C. per column: bucket offsets, and consecutive buckets packed greedily into windows of at
//      most 2,048 values -- one CTA's sort
//   D. scatter (bit pattern, row) into the column's bucketed order
//   E. per window: CTA radix sort in shared memory over the bits that differ inside the window,
//      raw_k = p_(k) / (k / n) in place, window minimum
//   F. per column: exclusive running minimum of the window minima from the right
//   G. per window: running minimum from the right seeded with (F), clip, scatter by row into pT
//   6. transpose pT into out
// A non-equality bucket above 2,048 values (never seen: the expected maximum is ~2x the mean of 512)
// raises a device flag; the call then repeats with the global sort, so the result never depends
// on the sampling.  Column mode therefore synchronises the stream once before it returns.
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
            if (keys) {
                keys[lin] = v;
                vals[lin] = (uint32_t)lin;
            }
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


// ---- per-column sample sort ---------------------------------------------------------------------
namespace colsort {

constexpr int kCap = 2048;                 // most values one CTA sorts at once
constexpr int kSortThreads = 256, kSortItems = kCap / kSortThreads;
constexpr int kChunkThreads = 512, kChunkItems = 16, kChunk = kChunkThreads * kChunkItems;
constexpr int kSampleThreads = 1024, kSampleItems = 8, kMaxSamples = kSampleThreads * kSampleItems;
constexpr int kMaxBuckets = 1024, kMinBuckets = 32;
constexpr int64_t kMinRows = 8192, kMaxRows = int64_t(1) << 19;
// cs_sort: radix passes (6 bits each) cover the top 12 differing bits of a window, the rest is
// fixed up.  Measured on 200,000 x 2,016 (ms): all bits 17.9; top 24 / 18 / 12 bits 10.7 / 9.3 / 8.4;
// 5-bit digits over 20 / 15 bits 9.3 / 8.4; 4-bit digits over 16 / 12 bits 8.9 / 8.4.
constexpr int kRadixBits = 6, kPartialBits = 12;
constexpr int kMaxFix = 6;
constexpr uint32_t kNoSort = 0x80000000u;

// order-preserving map double -> uint64 (the one CUB's radix sort uses); the all-ones pattern is
// kept free so that it can never tie with a padding key
__device__ __forceinline__ uint64_t twiddle(double x)
{
    const uint64_t b = (uint64_t)__double_as_longlong(x);
    const uint64_t k = b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
    return k == ~0ull ? ~0ull - 1 : k;
}
__device__ __forceinline__ double untwiddle(uint64_t k)
{
    const uint64_t b = k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull);
    return __longlong_as_double((long long)b);
}

struct Plan {
    int64_t n_rows, n_cols;
    int buckets, slots, samples, max_win;      // slots = 2 * buckets - 1
    const double *p_t;                         // [cols][rows]
    double *out_t;                             // adjusted values, same layout (may alias p_t)
    uint64_t *key;                             // [cols][rows] bucketed bit patterns, then raw values in place
    uint32_t *row;                             // [cols][rows] row of every bucketed value
    uint16_t *slot;                            // [cols][rows] bucket slot of every value (B -> D)
    uint64_t *splitters;                       // [cols][buckets - 1]
    uint32_t *count, *cursor;                  // [cols][slots]
    uint32_t *win_lo, *win_hi;                 // [cols][max_win] offsets inside the column; win_hi | kNoSort
    double *win_min, *win_carry;               // [cols][max_win]
    uint32_t *n_win;                           // [cols]
    uint32_t *fail;                            // [1]
};

__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

// A. splitters of one column
using SampleSort = cub::BlockRadixSort<uint64_t, kSampleThreads, kSampleItems>;
__global__ void __launch_bounds__(kSampleThreads) cs_splitters(const Plan q)
{
    using Sort = SampleSort;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    typename Sort::TempStorage &tmp = *reinterpret_cast<typename Sort::TempStorage *>(dyn_smem);
    const int64_t col = blockIdx.x, n = q.n_rows;
    const double *src = q.p_t + col * n;
    const int64_t stride = n / q.samples;                       // >= 1: rows >= kMinRows >= samples
    uint64_t keys[kSampleItems];
#pragma unroll
    for (int e = 0; e < kSampleItems; ++e) {
        const int s = threadIdx.x * kSampleItems + e;
        keys[e] = ~0ull;
        if (s < q.samples) {
            const int64_t pos = (int64_t)s * stride + mix32((uint32_t)s * 2654435761u + (uint32_t)col) % (uint32_t)stride;
            keys[e] = twiddle(__ldg(src + pos));
        }
    }
    Sort(tmp).Sort(keys);
    const int over = q.samples / q.buckets;                    // samples per bucket
#pragma unroll
    for (int e = 0; e < kSampleItems; ++e) {
        const int s = threadIdx.x * kSampleItems + e;
        if (s > 0 && s < q.samples && s % over == 0) q.splitters[col * (q.buckets - 1) + s / over - 1] = keys[e];
    }
    for (int i = threadIdx.x; i < q.slots; i += kSampleThreads) q.count[col * q.slots + i] = 0;
    if (threadIdx.x == 0 && col == 0) *q.fail = 0;
}

// slot of a key: 2 j for the open interval below splitter j, 2 j + 1 for "equal to splitter j".
// The binary search runs on the splitters' high words (32-bit shared loads, a third of the bank
// traffic of 64-bit ones); the few splitters that share the key's high word are then walked with
// full comparisons.
__device__ __forceinline__ int slot_of(const uint64_t *spl, const uint32_t *spl_hi, int buckets, uint64_t key)
{
    const uint32_t key_hi = (uint32_t)(key >> 32);
    int j = 0;
    for (int step = buckets >> 1; step >= 1; step >>= 1)
        if (spl_hi[j + step - 1] < key_hi) j += step;          // j = #splitters with a smaller high word
    while (j < buckets - 1 && spl[j] < key) ++j;               // same high word, smaller low word
    const bool eq = j < buckets - 1 && spl[j] == key;
    return 2 * j + (eq ? 1 : 0);
}

// B. counts per (column, slot);  D. scatter into bucketed order (kScatter)
template <bool kScatter>
__global__ void __launch_bounds__(kChunkThreads) cs_chunks(const Plan q)
{
    __shared__ uint64_t spl[kMaxBuckets];
    __shared__ uint32_t spl_hi[kMaxBuckets];
    __shared__ uint32_t hist[2 * kMaxBuckets];
    const int64_t col = blockIdx.y, n = q.n_rows;
    const int64_t lo = (int64_t)blockIdx.x * kChunk;
    if constexpr (!kScatter) {
        for (int i = threadIdx.x; i < q.buckets - 1; i += kChunkThreads) {
            const uint64_t v = q.splitters[col * (q.buckets - 1) + i];
            spl[i] = v;
            spl_hi[i] = (uint32_t)(v >> 32);
        }
    }
    for (int i = threadIdx.x; i < q.slots; i += kChunkThreads) hist[i] = 0;
    __syncthreads();
    const double *src = q.p_t + col * n;
    uint16_t *slots = q.slot + col * n;
    uint64_t key[kChunkItems];
    uint32_t where[kChunkItems];             // slot << 16 | rank inside this chunk's share of the slot
#pragma unroll
    for (int e = 0; e < kChunkItems; ++e) {
        const int64_t r = lo + threadIdx.x + e * kChunkThreads;
        key[e] = r < n ? twiddle(__ldcs(src + r)) : 0;
        if constexpr (kScatter) where[e] = r < n ? (uint32_t)__ldcs(slots + r) : 0u;
    }
#pragma unroll
    for (int e = 0; e < kChunkItems; ++e) {
        const int64_t r = lo + threadIdx.x + e * kChunkThreads;
        if (r < n) {
            int slot;
            if constexpr (kScatter) slot = (int)where[e];
            else slot = slot_of(spl, spl_hi, q.buckets, key[e]);
            const uint32_t rank = atomicAdd(&hist[slot], 1u);
            where[e] = ((uint32_t)slot << 16) | rank;         // rank < kChunk = 8192 < 2^16, slot < 2^11
            if constexpr (!kScatter) slots[r] = (uint16_t)slot;
        }
    }
    __syncthreads();
    if constexpr (!kScatter) {
        for (int i = threadIdx.x; i < q.slots; i += kChunkThreads)
            if (hist[i]) atomicAdd(&q.count[col * q.slots + i], hist[i]);
    } else {
        for (int i = threadIdx.x; i < q.slots; i += kChunkThreads)
            hist[i] = hist[i] ? atomicAdd(&q.cursor[col * q.slots + i], hist[i]) : 0u;   // becomes the base offset
        __syncthreads();
        uint64_t *dst_key = q.key + col * n;
        uint32_t *dst_row = q.row + col * n;
#pragma unroll
        for (int e = 0; e < kChunkItems; ++e) {
            const int64_t r = lo + threadIdx.x + e * kChunkThreads;
            if (r < n) {
                const uint32_t at = hist[where[e] >> 16] + (where[e] & 0xFFFFu);
                dst_key[at] = key[e];
                dst_row[at] = (uint32_t)r;
            }
        }
    }
}

// C. per column: slot offsets, windows
__global__ void __launch_bounds__(256) cs_plan(const Plan q)
{
    using Scan = cub::BlockScan<uint32_t, 256>;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ uint32_t start[2 * kMaxBuckets + 1];
    const int64_t col = blockIdx.x;
    constexpr int kPer = 2 * kMaxBuckets / 256;                 // 8 slots per thread
    uint32_t c[kPer], off[kPer];
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
        const int i = threadIdx.x * kPer + e;
        c[e] = i < q.slots ? q.count[col * q.slots + i] : 0u;
    }
    Scan(tmp).ExclusiveSum(c, off);
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
        const int i = threadIdx.x * kPer + e;
        if (i < q.slots) {
            start[i] = off[e];
            q.cursor[col * q.slots + i] = off[e];
        }
    }
    if (threadIdx.x == 0) start[q.slots] = (uint32_t)q.n_rows;
    __syncthreads();
    if (threadIdx.x != 0) return;
    // greedy packing of consecutive slots into windows of at most kCap values
    uint32_t *lo = q.win_lo + col * q.max_win, *hi = q.win_hi + col * q.max_win;
    uint32_t n_win = 0, cur_lo = 0, cur_hi = 0;
    for (int i = 0; i < q.slots; ++i) {
        const uint32_t b = start[i], e = start[i + 1], sz = e - b;
        if (sz == 0) continue;
        if (sz > (uint32_t)kCap) {
            if (cur_hi > cur_lo) { lo[n_win] = cur_lo; hi[n_win] = cur_hi; ++n_win; }
            if (!(i & 1)) atomicExch(q.fail, 1u);               // a range bucket too large to sort: redo globally
            for (uint32_t a = b; a < e; a += kCap) { lo[n_win] = a; hi[n_win] = min(e, a + kCap) | kNoSort; ++n_win; }
            cur_lo = cur_hi = e;
        } else if (cur_hi - cur_lo + sz > (uint32_t)kCap) {
            lo[n_win] = cur_lo; hi[n_win] = cur_hi; ++n_win;
            cur_lo = b; cur_hi = e;
        } else {
            if (cur_hi == cur_lo) cur_lo = b;
            cur_hi = e;
        }
    }
    if (cur_hi > cur_lo) { lo[n_win] = cur_lo; hi[n_win] = cur_hi; ++n_win; }
    q.n_win[col] = n_win;
}

// E. sort one window, raw values in place, window minimum
template <int kRadixBits, int kPartialBits>
__global__ void __launch_bounds__(kSortThreads, 3) cs_sort(const Plan q)
{
    using Sort = cub::BlockRadixSort<uint64_t, kSortThreads, kSortItems, uint32_t, kRadixBits>;
    using Reduce64 = cub::BlockReduce<uint64_t, kSortThreads>;
    using ReduceD = cub::BlockReduce<double, kSortThreads>;
    __shared__ union {
        typename Sort::TempStorage sort;
        typename Reduce64::TempStorage r64;
        typename ReduceD::TempStorage rd;
    } tmp;
    __shared__ uint64_t s_min, s_max;
    __shared__ uint64_t s_edge_k[2 * kSortThreads];
    __shared__ uint32_t s_edge_v[2 * kSortThreads];
    const int64_t col = blockIdx.y, n = q.n_rows;
    const uint32_t w = blockIdx.x;
    if (w >= q.n_win[col]) return;
    const uint32_t lo = q.win_lo[col * q.max_win + w], hi_raw = q.win_hi[col * q.max_win + w];
    const uint32_t hi = hi_raw & ~kNoSort, size = hi - lo;
    uint64_t *key = q.key + col * n + lo;
    uint32_t *row = q.row + col * n + lo;
    uint64_t k[kSortItems];
    uint32_t v[kSortItems];
    // blocked arrangement: item e of thread t is position t * kSortItems + e
#pragma unroll
    for (int e = 0; e < kSortItems; ++e) {
        const uint32_t i = threadIdx.x * kSortItems + e;
        k[e] = i < size ? key[i] : 0;
        v[e] = i < size ? row[i] : 0u;
    }
    if (!(hi_raw & kNoSort)) {
        uint64_t mn = ~0ull, mx = 0;
#pragma unroll
        for (int e = 0; e < kSortItems; ++e)
            if (threadIdx.x * kSortItems + e < size) { mn = min(mn, k[e]); mx = max(mx, k[e]); }
        mn = Reduce64(tmp.r64).Reduce(mn, cub::Min());
        __syncthreads();
        mx = Reduce64(tmp.r64).Reduce(mx, cub::Max());
        if (threadIdx.x == 0) { s_min = mn; s_max = mx; }
        __syncthreads();
        mn = s_min;
        const uint64_t range = s_max - mn;                       // < 2^64 - 1: the all-ones key is never produced
        if (range) {
            // keys relative to the window minimum; padding = range + 1 sorts after every real key
#pragma unroll
            for (int e = 0; e < kSortItems; ++e) k[e] = threadIdx.x * kSortItems + e < size ? k[e] - mn : range + 1;
            const int bits = 64 - __clzll((long long)(range + 1));
            // Radix passes over the top kPartialBits of the differing bits only: the values of a
            // window spread over its range, so few pairs agree in all of those, and what is left
            // is a handful of neighbours out of order -- repaired by odd-even transposition
            // passes on the full keys (a window that stays unsorted is sorted over all bits).
            const int low = bits > kPartialBits + kRadixBits ? bits - kPartialBits : 0;
            Sort(tmp.sort).Sort(k, v, low, bits);
            __syncthreads();
            if (low) {
                bool sorted = false;
                for (int it = 0; it < kMaxFix && !sorted; ++it) {
                    int swapped = 0;
#pragma unroll
                    for (int e = 0; e + 1 < kSortItems; e += 2)            // pairs (2i, 2i + 1): inside the thread
                        if (k[e + 1] < k[e]) {
                            const uint64_t tk = k[e]; k[e] = k[e + 1]; k[e + 1] = tk;
                            const uint32_t tv = v[e]; v[e] = v[e + 1]; v[e + 1] = tv;
                            swapped = 1;
                        }
#pragma unroll
                    for (int e = 1; e + 1 < kSortItems; e += 2)            // pairs (2i + 1, 2i + 2) inside the thread
                        if (k[e + 1] < k[e]) {
                            const uint64_t tk = k[e]; k[e] = k[e + 1]; k[e + 1] = tk;
                            const uint32_t tv = v[e]; v[e] = v[e + 1]; v[e + 1] = tv;
                            swapped = 1;
                        }
                    // ... and the pair across the thread boundary: last item of t with first of t + 1
                    s_edge_k[2 * threadIdx.x] = k[0]; s_edge_v[2 * threadIdx.x] = v[0];
                    s_edge_k[2 * threadIdx.x + 1] = k[kSortItems - 1]; s_edge_v[2 * threadIdx.x + 1] = v[kSortItems - 1];
                    __syncthreads();
                    if (threadIdx.x + 1 < kSortThreads) {
                        const uint64_t nk = s_edge_k[2 * threadIdx.x + 2];
                        if (nk < k[kSortItems - 1]) { k[kSortItems - 1] = nk; v[kSortItems - 1] = s_edge_v[2 * threadIdx.x + 2]; swapped = 1; }
                    }
                    if (threadIdx.x > 0) {
                        const uint64_t pk = s_edge_k[2 * threadIdx.x - 1];
                        if (k[0] < pk) { k[0] = pk; v[0] = s_edge_v[2 * threadIdx.x - 1]; swapped = 1; }
                    }
                    sorted = !__syncthreads_or(swapped);
                }
                if (!sorted) {
                    Sort(tmp.sort).Sort(k, v, 0, bits);
                    __syncthreads();
                }
            }
#pragma unroll
            for (int e = 0; e < kSortItems; ++e) k[e] += mn;
        }
    }
    const double len = (double)n;
    double m = pos_inf();
#pragma unroll
    for (int e = 0; e < kSortItems; ++e) {
        const uint32_t i = threadIdx.x * kSortItems + e;
        if (i < size) {
            const double ecdf = __ddiv_rn((double)(lo + i + 1), len);
            const double raw = __ddiv_rn(untwiddle(k[e]), ecdf);
            key[i] = (uint64_t)__double_as_longlong(raw);
            row[i] = v[e];
            m = NanMin()(m, raw);
        }
    }
    m = ReduceD(tmp.rd).Reduce(m, NanMin());
    if (threadIdx.x == 0) q.win_min[col * q.max_win + w] = m;
}

// F. one warp per column: win_carry[w] = min over the windows to the right of w
__global__ void __launch_bounds__(256) cs_carry(const Plan q)
{
    const int64_t col = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (col >= q.n_cols) return;
    const int lane = threadIdx.x & 31;
    const int n_win = (int)q.n_win[col];
    const double *mn = q.win_min + col * q.max_win;
    double *carry = q.win_carry + col * q.max_win;
    double running = pos_inf();
    for (int base = n_win - 1; base >= 0; base -= 32) {
        const int w = base - lane;                               // lane 0 is the right-most window of this batch
        double v = w >= 0 ? mn[w] : pos_inf();
        double incl = v;                                         // inclusive scan over lanes 0..lane
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl = NanMin()(o, incl);
        }
        double excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
        if (lane == 0) excl = pos_inf();
        if (w >= 0) carry[w] = NanMin()(running, excl);
        running = NanMin()(running, __shfl_sync(0xFFFFFFFFu, incl, 31));
    }
}

// G. running minimum from the right inside the window, seeded with the carry; clip; scatter by row
__global__ void __launch_bounds__(kSortThreads) cs_finish(const Plan q)
{
    using Scan = cub::BlockScan<double, kSortThreads>;
    __shared__ typename Scan::TempStorage tmp;
    const int64_t col = blockIdx.y, n = q.n_rows;
    const uint32_t w = blockIdx.x;
    if (w >= q.n_win[col]) return;
    const uint32_t lo = q.win_lo[col * q.max_win + w], hi = q.win_hi[col * q.max_win + w] & ~kNoSort;
    const double *raw = reinterpret_cast<const double *>(q.key + col * n);
    const uint32_t *row = q.row + col * n;
    double *out = q.out_t + col * n;
    const int64_t first = (int64_t)hi - 1 - (int64_t)threadIdx.x * kSortItems;   // this thread's right-most slot
    double v[kSortItems];
    uint32_t id[kSortItems];
#pragma unroll
    for (int e = 0; e < kSortItems; ++e) {
        const int64_t s = first - e;
        v[e] = s >= (int64_t)lo ? __ldcs(raw + s) : pos_inf();
        id[e] = s >= (int64_t)lo ? __ldcs(row + s) : 0u;
    }
    double run = pos_inf();
#pragma unroll
    for (int e = 0; e < kSortItems; ++e) {
        run = NanMin()(run, v[e]);
        v[e] = run;
    }
    double before;
    Scan(tmp).ExclusiveScan(run, before, pos_inf(), NanMin());
    before = NanMin()(before, q.win_carry[col * q.max_win + w]);
#pragma unroll
    for (int e = 0; e < kSortItems; ++e) {
        if (first - e >= (int64_t)lo) {
            double a = NanMin()(before, v[e]);
            if (a > 1.0) a = 1.0;
            out[id[e]] = a;
        }
    }
}

inline int buckets_for(int64_t n_rows)
{
    int b = kMinBuckets;
    while (b < kMaxBuckets && (int64_t)b * 512 < n_rows) b <<= 1;
    return b;
}
inline bool applies(int64_t n_rows, int64_t n_cols)
{
    return n_cols >= 1 && n_rows >= kMinRows && n_rows <= kMaxRows && !getenv("SD_BH_GLOBAL_SORT");
}
inline int max_windows(int64_t n_rows) { return (int)(4 * n_rows / kCap + 4); }
// bytes of the small per-column tables (carved out of the second key buffer of the global path)
inline size_t tables_bytes(int64_t n_rows, int64_t n_cols)
{
    const size_t b = (size_t)buckets_for(n_rows), w = (size_t)max_windows(n_rows);
    return (size_t)n_cols * ((b - 1) * 8 + (2 * b - 1) * 4 * 2 + w * (4 + 4 + 8 + 8) + 4) + 64 * 16;
}

}  // namespace colsort

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
    if (by_column) SD_REQUIRE(tiles.y <= 65535u, "sd_bh_adjust: more than 2,097,120 columns per call (split the columns)");

    // ---- column mode on a real matrix: sample sort per column (namespace colsort) ----------------
    namespace cs = sd::colsort;
    if (by_column && cs::applies(n_rows, n_cols) && n_cols <= 65535 &&
        cs::tables_bytes(n_rows, n_cols) <= (size_t)n * 8) {
        cs::Plan q{};
        q.n_rows = n_rows; q.n_cols = n_cols;
        q.buckets = cs::buckets_for(n_rows);
        q.slots = 2 * q.buckets - 1;
        q.samples = std::min(cs::kMaxSamples, q.buckets * 16);
        q.max_win = cs::max_windows(n_rows);
        q.p_t = p_t; q.out_t = p_t;
        q.key = reinterpret_cast<uint64_t *>(key_a);
        q.row = val_a;
        q.slot = reinterpret_cast<uint16_t *>(val_b);
        // the small per-column tables live in the global path's second key buffer
        char *t = reinterpret_cast<char *>(key_b);
        auto carve = [&](size_t bytes) { char *r = t; t += (bytes + 15) & ~(size_t)15; return r; };
        q.splitters = reinterpret_cast<uint64_t *>(carve((size_t)n_cols * (q.buckets - 1) * 8));
        q.win_min = reinterpret_cast<double *>(carve((size_t)n_cols * q.max_win * 8));
        q.win_carry = reinterpret_cast<double *>(carve((size_t)n_cols * q.max_win * 8));
        q.count = reinterpret_cast<uint32_t *>(carve((size_t)n_cols * q.slots * 4));
        q.cursor = reinterpret_cast<uint32_t *>(carve((size_t)n_cols * q.slots * 4));
        q.win_lo = reinterpret_cast<uint32_t *>(carve((size_t)n_cols * q.max_win * 4));
        q.win_hi = reinterpret_cast<uint32_t *>(carve((size_t)n_cols * q.max_win * 4));
        q.n_win = reinterpret_cast<uint32_t *>(carve((size_t)n_cols * 4));
        q.fail = reinterpret_cast<uint32_t *>(carve(16));
        // key_a doubles as the transpose's (unused here) key output only in the global path: write pT alone
        sd::bh_transpose_in<<<tiles, 256, 0, stream>>>(n_rows, n_cols, p, ld_p, p_t, nullptr, nullptr);
        if (int rc = sd::check_launch("bh_transpose_in")) return rc;
        const dim3 chunks((unsigned)((n_rows + cs::kChunk - 1) / cs::kChunk), (unsigned)n_cols);
        const dim3 windows((unsigned)q.max_win, (unsigned)n_cols);
        constexpr size_t kSplitSmem = sizeof(cs::SampleSort::TempStorage);
        SD_CHECK_CUDA(cudaFuncSetAttribute(cs::cs_splitters, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSplitSmem));
        cs::cs_splitters<<<(unsigned)n_cols, cs::kSampleThreads, kSplitSmem, stream>>>(q);
        if (int rc = sd::check_launch("cs_splitters")) return rc;
        cs::cs_chunks<false><<<chunks, cs::kChunkThreads, 0, stream>>>(q);
        if (int rc = sd::check_launch("cs_chunks<count>")) return rc;
        cs::cs_plan<<<(unsigned)n_cols, 256, 0, stream>>>(q);
        if (int rc = sd::check_launch("cs_plan")) return rc;
        cs::cs_chunks<true><<<chunks, cs::kChunkThreads, 0, stream>>>(q);
        if (int rc = sd::check_launch("cs_chunks<scatter>")) return rc;
        cs::cs_sort<cs::kRadixBits, cs::kPartialBits><<<windows, cs::kSortThreads, 0, stream>>>(q);
        if (int rc = sd::check_launch("cs_sort")) return rc;
        cs::cs_carry<<<(unsigned)((n_cols + 7) / 8), 256, 0, stream>>>(q);
        if (int rc = sd::check_launch("cs_carry")) return rc;
        cs::cs_finish<<<windows, cs::kSortThreads, 0, stream>>>(q);
        if (int rc = sd::check_launch("cs_finish")) return rc;
        sd::bh_transpose_out<<<tiles, 256, 0, stream>>>(n_rows, n_cols, p_t, out, ld_out);
        if (int rc = sd::check_launch("bh_transpose_out")) return rc;
        uint32_t failed = 0;
        SD_CHECK_CUDA(cudaMemcpyAsync(&failed, q.fail, sizeof failed, cudaMemcpyDeviceToHost, stream));
        SD_CHECK_CUDA(cudaStreamSynchronize(stream));
        if (getenv("SD_BH_TEST_FAIL")) failed = 1;          // tests: exercise the fall-through
        if (!failed) return SD_OK;
        // a bucket too large to sort in one CTA: fall through to the global sort (p is intact unless
        // the call is in place, in which case `out` == p was overwritten -- pT still holds ... no:
        // pT was overwritten by the scatter too, so an in-place call cannot be repeated)
        if (p == out)
            return sd::fail(SD_ERR_UNSUPPORTED, "sd_bh_adjust: degenerate value distribution in an in-place call; "
                                                "call again with a separate output or SD_BH_GLOBAL_SORT=1");
    }

    if (by_column) {
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
