// sd_hostio.cpp -- host-side text formatting of the result matrices (no CUDA).
//
// After the kernels, the reference's writers dominate wall time: one python f-string per cell
// (SPLICEDICE.py:332-353 `f'{x:.0f}'` / `f'{x:.3f}'`, counts_to_ps.py:69 `f"{x:0.3f}"`), 4e8 of
// them at 400k x 1,000.  sd_host_format_rows writes whole rows -- "name<TAB>v<TAB>v...\n" --
// byte-identical to those f-strings, on all host threads.
//
// Exactness.  python formats the exact binary value with round-half-even.  A float32 times 1000
// is exact in binary64 (24 + 10 bits), so nearbyint() under the default rounding mode is already
// the correctly rounded scaled integer; for float64 the product is split exactly into p + e
// with fma() and the half-way cases of p are settled by the sign of e.  Magnitudes beyond 2^52/1000, infinities and NaN take snprintf / literals
// (python prints "nan" for either NaN sign, "inf"/"-inf").
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <charconv>
#include <string>
#include <thread>
#include <vector>

#include "sd_common.cuh"

namespace {

// Bounded output cursor over caller memory.  Writes past `end` are dropped and counted, so a
// too-small buffer yields the exact size needed instead of an overrun.
struct Sink {
    char *p, *end;
    size_t dropped = 0;
    Sink(char *begin, char *end_) : p(begin), end(end_) {}
    void push_back(char c)
    {
        if (p < end) *p++ = c;
        else ++dropped;
    }
    void append(const char *q, size_t n)
    {
        if ((size_t)(end - p) >= n) { memcpy(p, q, n); p += n; }
        else dropped += n;
    }
    void fill(size_t n, char c)
    {
        if ((size_t)(end - p) >= n) { memset(p, c, n); p += n; }
        else dropped += n;
    }
};

inline int uint_digits(char *buf, uint64_t v)      // decimal digits of v, most significant first
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < n; ++i) buf[i] = tmp[n - 1 - i];
    return n;
}

inline void put_uint(Sink &s, uint64_t v)
{
    char buf[24];
    s.append(buf, (size_t)uint_digits(buf, v));
}

// %.3f of a finite double whose scaled value fits the exact path
inline bool put_fixed3(Sink &s, double v)
{
    const double av = fabs(v);
    if (!(av < 4.0e12)) return false;
    // av * 1000 = p + e exactly (two-product); r = nearest-even integer of p; the true residual
    // is (p - r) + e with p - r exact, and only an exact half in p - r can be tipped by e
    const double p = av * 1000.0;
    const double e = fma(av, 1000.0, -p);
    double r = nearbyint(p);
    const double t = p - r;
    if (t == 0.5 && e > 0.0) r += 1.0;
    else if (t == -0.5 && e < 0.0) r -= 1.0;
    const uint64_t q = (uint64_t)r;
    char buf[32];
    int n = 0;
    if (signbit(v)) buf[n++] = '-';
    n += uint_digits(buf + n, q / 1000);
    const unsigned f = (unsigned)(q % 1000);
    buf[n++] = '.';
    buf[n++] = (char)('0' + f / 100);
    buf[n++] = (char)('0' + f / 10 % 10);
    buf[n++] = (char)('0' + f % 10);
    s.append(buf, (size_t)n);
    return true;
}

inline void put_value_f(Sink &s, double v)
{
    if (isnan(v)) { s.append("nan", 3); return; }
    if (isinf(v)) { if (v < 0) s.append("-inf", 4); else s.append("inf", 3); return; }
    if (put_fixed3(s, v)) return;
    char buf[400];
    int n = snprintf(buf, sizeof buf, "%.3f", v);
    s.append(buf, (size_t)n);
}

// python's repr(float) / str(numpy.float64): shortest digits that round-trip, fixed notation when
// the decimal exponent is in [-4, 16), otherwise d.ddde[+-]XX with at least two exponent digits
inline void put_repr(Sink &s, double v)
{
    if (isnan(v)) { s.append("nan", 3); return; }
    if (isinf(v)) { if (v < 0) s.append("-inf", 4); else s.append("inf", 3); return; }
    if (v == 0.0) { if (signbit(v)) s.append("-0.0", 4); else s.append("0.0", 3); return; }
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest
    const char *p = buf, *end = res.ptr;
    char out[64];
    int n = 0;
    if (*p == '-') { out[n++] = '-'; ++p; }
    char digits[40];
    int nd = 0;
    const char *e = p;
    for (; e < end && *e != 'e'; ++e)
        if (*e != '.') digits[nd++] = *e;
    int exp10 = 0;
    {
        const char *q = e + 1;
        const bool neg = *q == '-';
        if (*q == '-' || *q == '+') ++q;
        for (; q < end; ++q) exp10 = exp10 * 10 + (*q - '0');
        if (neg) exp10 = -exp10;
    }
    if (exp10 >= -4 && exp10 < 16) {
        if (exp10 < 0) {
            out[n++] = '0'; out[n++] = '.';
            for (int i = 0; i < -exp10 - 1; ++i) out[n++] = '0';
            memcpy(out + n, digits, (size_t)nd); n += nd;
        } else {
            const int int_digits = exp10 + 1;
            if (nd <= int_digits) {
                memcpy(out + n, digits, (size_t)nd); n += nd;
                for (int i = 0; i < int_digits - nd; ++i) out[n++] = '0';
                out[n++] = '.'; out[n++] = '0';
            } else {
                memcpy(out + n, digits, (size_t)int_digits); n += int_digits;
                out[n++] = '.';
                memcpy(out + n, digits + int_digits, (size_t)(nd - int_digits)); n += nd - int_digits;
            }
        }
    } else {
        out[n++] = digits[0];
        if (nd > 1) { out[n++] = '.'; memcpy(out + n, digits + 1, (size_t)(nd - 1)); n += nd - 1; }
        out[n++] = 'e';
        out[n++] = exp10 < 0 ? '-' : '+';
        const int ae = exp10 < 0 ? -exp10 : exp10;
        if (ae < 10) out[n++] = '0';
        n += uint_digits(out + n, (uint64_t)ae);
    }
    s.append(out, (size_t)n);
}

template <class Get>
void format_range(int64_t r0, int64_t r1, int32_t cols, const char *names, const int64_t *name_off, Get get, Sink *out)
{
    for (int64_t r = r0; r < r1; ++r) {
        if (names) {       // f"{name}\t{tab.join(values)}\n": the tab is there even without values
            out->append(names + name_off[r], (size_t)(name_off[r + 1] - name_off[r]));
            out->push_back('\t');
        }
        for (int32_t c = 0; c < cols; ++c) {
            if (c) out->push_back('\t');
            get(*out, r, c);
        }
        out->push_back('\n');
    }
}

// Formats rows [0, rows) on n_threads threads, thread t writing its row range into the slice
// [cap * r0 / rows, cap * r1 / rows) of `out`.  seg_off / seg_len receive each thread's text;
// returns the number of bytes that did not fit (0 = complete).
size_t format_segments(int kind, const void *matrix, int64_t rows, int32_t cols, int64_t ld, const char *names,
                       const int64_t *name_off, char *out, size_t cap, int n_threads, size_t *seg_off, size_t *seg_len)
{
    std::vector<size_t> dropped((size_t)n_threads, 0);
    auto work = [&](int t) {
        const int64_t r0 = rows * t / n_threads, r1 = rows * (t + 1) / n_threads;
        const size_t b0 = (size_t)((long double)cap * r0 / rows), b1 = (size_t)((long double)cap * r1 / rows);
        Sink sink(out ? out + b0 : nullptr, out ? out + b1 : nullptr);
        if (kind == 0) {
            const float *m = static_cast<const float *>(matrix);
            format_range(r0, r1, cols, names, name_off,
                         [m, ld](Sink &s, int64_t r, int32_t c) { put_value_f(s, (double)m[r * ld + c]); }, &sink);
        } else if (kind == 1) {
            const double *m = static_cast<const double *>(matrix);
            format_range(r0, r1, cols, names, name_off,
                         [m, ld](Sink &s, int64_t r, int32_t c) { put_value_f(s, m[r * ld + c]); }, &sink);
        } else if (kind == 3) {
            const double *m = static_cast<const double *>(matrix);
            format_range(r0, r1, cols, names, name_off,
                         [m, ld](Sink &s, int64_t r, int32_t c) { put_repr(s, m[r * ld + c]); }, &sink);
        } else {
            const int32_t *m = static_cast<const int32_t *>(matrix);
            format_range(r0, r1, cols, names, name_off,
                         [m, ld](Sink &s, int64_t r, int32_t c) {
                             const int64_t v = m[r * ld + c];
                             if (v < 0) s.push_back('-');
                             put_uint(s, (uint64_t)(v < 0 ? -v : v));
                         },
                         &sink);
        }
        seg_off[t] = b0;
        seg_len[t] = out ? (size_t)(sink.p - (out + b0)) : 0;
        dropped[t] = sink.dropped;
    };
    if (n_threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t) pool.emplace_back(work, t);
        for (auto &th : pool) th.join();
    }
    size_t lost = 0;
    for (size_t d : dropped) lost += d;
    return lost;
}

int pick_threads(int n_threads, int64_t rows)
{
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    return (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, rows / 64));
}

}  // namespace

extern "C" int sd_host_format_rows(int kind, const void *matrix, int64_t rows, int32_t cols, int64_t ld,
                                   const char *names, const int64_t *name_off, char *out, size_t cap,
                                   size_t *written, int n_threads)
{
    SD_REQUIRE(kind >= 0 && kind <= 3,
               "sd_host_format_rows: kind must be 0 (f32 %%.3f), 1 (f64 %%.3f), 2 (i32) or 3 (f64 repr)");
    SD_REQUIRE(rows >= 0 && cols >= 0 && ld >= cols && written, "sd_host_format_rows: bad shape");
    SD_REQUIRE((rows == 0 || cols == 0 || matrix) && (!names || name_off), "sd_host_format_rows: null pointer");
    *written = 0;
    if (rows == 0) return SD_OK;
    n_threads = pick_threads(n_threads, rows);
    std::vector<size_t> off((size_t)n_threads), len((size_t)n_threads);
    const size_t lost = format_segments(kind, matrix, rows, cols, ld, names, name_off, out, out ? cap : 0, n_threads,
                                        off.data(), len.data());
    size_t total = 0;
    for (size_t l : len) total += l;
    if (lost) {
        // the slices are proportional to the row counts, so ask for the worst slice's ratio everywhere
        *written = (total + lost) * 2 + 4096;
        return sd::fail(SD_ERR_WORKSPACE, "sd_host_format_rows: about %zu bytes needed, %zu given", *written, cap);
    }
    // close the gaps between the threads' slices (every move goes towards the front, in order)
    size_t pos = 0;
    for (int t = 0; t < n_threads; ++t) {
        if (off[t] != pos) memmove(out + pos, out + off[t], len[t]);
        pos += len[t];
    }
    *written = total;
    return SD_OK;
}

// As sd_host_format_rows, but the text stays where each thread wrote it: segment k is
// out[seg_off[k], seg_off[k] + seg_len[k]), k < *n_segments <= max_segments, to be written out in
// order (the writers stream row chunks to a file this way without compacting or copying).
extern "C" int sd_host_format_rows_segments(int kind, const void *matrix, int64_t rows, int32_t cols, int64_t ld,
                                            const char *names, const int64_t *name_off, char *out, size_t cap,
                                            int64_t *seg_off, int64_t *seg_len, int32_t max_segments,
                                            int32_t *n_segments, size_t *needed, int n_threads)
{
    SD_REQUIRE(kind >= 0 && kind <= 3, "sd_host_format_rows_segments: kind must be 0..3");
    SD_REQUIRE(rows >= 0 && cols >= 0 && ld >= cols && seg_off && seg_len && n_segments && needed && max_segments >= 1,
               "sd_host_format_rows_segments: bad arguments");
    SD_REQUIRE((rows == 0 || ((cols == 0 || matrix) && out)) && (!names || name_off),
               "sd_host_format_rows_segments: null pointer");
    *n_segments = 0;
    *needed = 0;
    if (rows == 0) return SD_OK;
    n_threads = std::min(pick_threads(n_threads, rows), (int)max_segments);
    std::vector<size_t> off((size_t)n_threads), len((size_t)n_threads);
    const size_t lost = format_segments(kind, matrix, rows, cols, ld, names, name_off, out, cap, n_threads, off.data(),
                                        len.data());
    if (lost) {
        size_t total = 0;
        for (size_t l : len) total += l;
        *needed = (total + lost) * 2 + 4096;
        return sd::fail(SD_ERR_WORKSPACE, "sd_host_format_rows_segments: about %zu bytes needed, %zu given", *needed, cap);
    }
    for (int t = 0; t < n_threads; ++t) {
        seg_off[t] = (int64_t)off[t];
        seg_len[t] = (int64_t)len[t];
    }
    *n_segments = n_threads;
    return SD_OK;
}

// ---- table reader -------------------------------------------------------------------------
// "header\nname<TAB>v<TAB>v...\n" files (the reference's _inclusionCounts.tsv and friends, read
// there one python float per cell: counts_to_ps.py:43-51, pairwise_fisher.py:46-61,
// ir_table.py:72-80).  The file is memory-mapped, line starts are indexed once, and the value
// fields are parsed on several threads into a dense float64 matrix.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

struct TableFile {
    const char *data = nullptr;
    size_t size = 0;
    int fd = -1;
    std::vector<size_t> line_start;     // data rows only (header excluded); one extra entry = end
    size_t header_end = 0;              // offset one past the header line's '\n' (or size)
    int32_t cols = 0;
    int64_t name_bytes = 0;
    ~TableFile()
    {
        if (data) munmap(const_cast<char *>(data), size);
        if (fd >= 0) close(fd);
    }
};

inline const char *rstrip(const char *p, const char *end)
{
    while (end > p && (end[-1] == ' ' || end[-1] == '\t' || end[-1] == '\r' || end[-1] == '\n' || end[-1] == '\f' ||
                       end[-1] == '\v'))
        --end;
    return end;
}

// python float(): plain digits fast, everything else through strtod
inline bool parse_value(const char *p, const char *e, double *out)
{
    if (p == e) return false;
    const char *q = p;
    uint64_t v = 0;
    int digits = 0;
    while (q < e && *q >= '0' && *q <= '9' && digits < 18) { v = v * 10 + (uint64_t)(*q - '0'); ++q; ++digits; }
    if (q == e && digits) { *out = (double)v; return true; }
    char buf[96];
    const size_t n = (size_t)(e - p);
    if (n >= sizeof buf) return false;
    memcpy(buf, p, n);
    buf[n] = 0;
    char *endp = nullptr;
    *out = strtod(buf, &endp);
    while (*endp == ' ') ++endp;
    return endp == buf + n && endp != buf;
}

}  // namespace

extern "C" {

void *sd_host_table_open(const char *path, int64_t *rows, int32_t *cols, int64_t *header_bytes, int64_t *name_bytes)
{
    if (!path || !rows || !cols || !header_bytes || !name_bytes) {
        sd::fail(SD_ERR_INVALID, "sd_host_table_open: null pointer");
        return nullptr;
    }
    TableFile *t = new TableFile();
    t->fd = ::open(path, O_RDONLY);
    struct stat st;
    if (t->fd < 0 || fstat(t->fd, &st) != 0) {
        sd::fail(SD_ERR_INVALID, "sd_host_table_open: cannot open %s", path);
        delete t;
        return nullptr;
    }
    t->size = (size_t)st.st_size;
    if (t->size) {
        void *p = mmap(nullptr, t->size, PROT_READ, MAP_PRIVATE, t->fd, 0);
        if (p == MAP_FAILED) {
            sd::fail(SD_ERR_INVALID, "sd_host_table_open: cannot map %s", path);
            delete t;
            return nullptr;
        }
        t->data = static_cast<const char *>(p);
    }
    const char *end = t->data + t->size;
    const char *nl = t->size ? (const char *)memchr(t->data, '\n', t->size) : nullptr;
    t->header_end = nl ? (size_t)(nl - t->data) + 1 : t->size;
    for (const char *p = t->data + t->header_end; p < end;) {
        t->line_start.push_back((size_t)(p - t->data));
        const char *n2 = (const char *)memchr(p, '\n', (size_t)(end - p));
        p = n2 ? n2 + 1 : end;
    }
    t->line_start.push_back(t->size);
    // columns from the first data row; names measured on every row
    const int64_t n_rows = (int64_t)t->line_start.size() - 1;
    for (int64_t r = 0; r < n_rows; ++r) {
        const char *p = t->data + t->line_start[(size_t)r];
        const char *e = rstrip(p, t->data + t->line_start[(size_t)r + 1]);
        const char *tab = (const char *)memchr(p, '\t', (size_t)(e - p));
        t->name_bytes += (tab ? tab : e) - p;
        if (r == 0) {
            int32_t c = 0;
            for (const char *q = p; q < e; ++q) c += *q == '\t';
            t->cols = c;
        }
    }
    *rows = n_rows;
    *cols = t->cols;
    *header_bytes = (int64_t)t->header_end;
    *name_bytes = t->name_bytes;
    return t;
}

void sd_host_table_close(void *handle) { delete static_cast<TableFile *>(handle); }

// header: the raw first line (header_bytes, newline included); names: concatenated row names with
// name_off[rows + 1]; values: float64 [rows, cols] with leading dimension ld.  A row whose field
// count differs from the first row's, or a field python's float() would reject, fails the call.
int sd_host_table_read(void *handle, char *header, char *names, int64_t *name_off, double *values, int64_t ld,
                       int32_t n_threads)
{
    SD_REQUIRE(handle && name_off, "sd_host_table_read: null pointer");
    TableFile *t = static_cast<TableFile *>(handle);
    const int64_t rows = (int64_t)t->line_start.size() - 1;
    SD_REQUIRE(rows == 0 || t->cols == 0 || (values && ld >= t->cols), "sd_host_table_read: bad value buffer");
    if (header && t->header_end) memcpy(header, t->data, t->header_end);
    name_off[0] = 0;
    for (int64_t r = 0; r < rows; ++r) {
        const char *p = t->data + t->line_start[(size_t)r];
        const char *e = rstrip(p, t->data + t->line_start[(size_t)r + 1]);
        const char *tab = (const char *)memchr(p, '\t', (size_t)(e - p));
        const int64_t n = (tab ? tab : e) - p;
        if (names) memcpy(names + name_off[r], p, (size_t)n);
        name_off[r + 1] = name_off[r] + n;
    }
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, rows / 256));
    std::vector<int64_t> bad((size_t)n_threads, -1);
    auto work = [&](int w) {
        const int64_t r0 = rows * w / n_threads, r1 = rows * (w + 1) / n_threads;
        for (int64_t r = r0; r < r1; ++r) {
            const char *p = t->data + t->line_start[(size_t)r];
            const char *e = rstrip(p, t->data + t->line_start[(size_t)r + 1]);
            const char *q = (const char *)memchr(p, '\t', (size_t)(e - p));
            int32_t c = 0;
            while (q) {
                const char *s = q + 1;
                const char *nx = (const char *)memchr(s, '\t', (size_t)(e - s));
                const char *fe = nx ? nx : e;
                if (c >= t->cols || !parse_value(s, fe, values + r * ld + c)) { bad[(size_t)w] = r; return; }
                ++c;
                q = nx;
            }
            if (c != t->cols) { bad[(size_t)w] = r; return; }
        }
    };
    if (n_threads == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (int w = 0; w < n_threads; ++w) pool.emplace_back(work, w);
        for (auto &th : pool) th.join();
    }
    for (int64_t b : bad)
        if (b >= 0)
            return sd::fail(SD_ERR_INVALID, "sd_host_table_read: data row %lld is ragged or holds a non-numeric field",
                            (long long)(b + 1));
    return SD_OK;
}

}  // extern "C"
