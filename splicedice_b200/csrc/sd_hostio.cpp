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

inline void put_uint(std::string &s, uint64_t v)
{
    char buf[24];
    int n = 0;
    do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) s.push_back(buf[--n]);
}

// %.3f of a finite double whose scaled value fits the exact path
inline bool put_fixed3(std::string &s, double v)
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
    if (signbit(v)) s.push_back('-');
    put_uint(s, q / 1000);
    const unsigned f = (unsigned)(q % 1000);
    s.push_back('.');
    s.push_back((char)('0' + f / 100));
    s.push_back((char)('0' + f / 10 % 10));
    s.push_back((char)('0' + f % 10));
    return true;
}

inline void put_value_f(std::string &s, double v)
{
    if (isnan(v)) { s += "nan"; return; }
    if (isinf(v)) { s += v < 0 ? "-inf" : "inf"; return; }
    if (put_fixed3(s, v)) return;
    char buf[400];
    int n = snprintf(buf, sizeof buf, "%.3f", v);
    s.append(buf, (size_t)n);
}

// python's repr(float) / str(numpy.float64): shortest digits that round-trip, fixed notation when
// the decimal exponent is in [-4, 16), otherwise d.ddde[+-]XX with at least two exponent digits
inline void put_repr(std::string &s, double v)
{
    if (isnan(v)) { s += "nan"; return; }
    if (isinf(v)) { s += v < 0 ? "-inf" : "inf"; return; }
    if (v == 0.0) { s += signbit(v) ? "-0.0" : "0.0"; return; }
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest
    *res.ptr = 0;
    const char *p = buf, *end = res.ptr;
    if (*p == '-') { s.push_back('-'); ++p; }
    const char *e = p;
    while (e < end && *e != 'e') ++e;
    char digits[40];
    int nd = 0;
    for (const char *q = p; q < e; ++q)
        if (*q != '.') digits[nd++] = *q;
    const int exp10 = atoi(e + 1);
    if (exp10 >= -4 && exp10 < 16) {
        if (exp10 < 0) {
            s += "0.";
            s.append((size_t)(-exp10 - 1), '0');
            s.append(digits, (size_t)nd);
        } else {
            const int int_digits = exp10 + 1;
            if (nd <= int_digits) {
                s.append(digits, (size_t)nd);
                s.append((size_t)(int_digits - nd), '0');
                s += ".0";
            } else {
                s.append(digits, (size_t)int_digits);
                s.push_back('.');
                s.append(digits + int_digits, (size_t)(nd - int_digits));
            }
        }
    } else {
        s.push_back(digits[0]);
        if (nd > 1) { s.push_back('.'); s.append(digits + 1, (size_t)(nd - 1)); }
        s.push_back('e');
        s.push_back(exp10 < 0 ? '-' : '+');
        const int ae = exp10 < 0 ? -exp10 : exp10;
        if (ae < 10) s.push_back('0');
        put_uint(s, (uint64_t)ae);
    }
}

template <class Get>
void format_range(int64_t r0, int64_t r1, int32_t cols, const char *names, const int64_t *name_off, Get get,
                  std::string *out)
{
    out->reserve((size_t)(r1 - r0) * ((size_t)cols * 8 + 32));
    for (int64_t r = r0; r < r1; ++r) {
        if (names) out->append(names + name_off[r], (size_t)(name_off[r + 1] - name_off[r]));
        for (int32_t c = 0; c < cols; ++c) {
            if (names || c) out->push_back('\t');
            get(*out, r, c);
        }
        out->push_back('\n');
    }
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
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, rows / 64));
    std::vector<std::string> parts((size_t)n_threads);
    auto work = [&](int t) {
        const int64_t r0 = rows * t / n_threads, r1 = rows * (t + 1) / n_threads;
        if (kind == 0) {
            const float *m = static_cast<const float *>(matrix);
            format_range(r0, r1, cols, names, name_off,
                         [m, ld](std::string &s, int64_t r, int32_t c) { put_value_f(s, (double)m[r * ld + c]); }, &parts[t]);
        } else if (kind == 1) {
            const double *m = static_cast<const double *>(matrix);
            format_range(r0, r1, cols, names, name_off,
                         [m, ld](std::string &s, int64_t r, int32_t c) { put_value_f(s, m[r * ld + c]); }, &parts[t]);
        } else if (kind == 3) {
            const double *m = static_cast<const double *>(matrix);
            format_range(r0, r1, cols, names, name_off,
                         [m, ld](std::string &s, int64_t r, int32_t c) { put_repr(s, m[r * ld + c]); }, &parts[t]);
        } else {
            const int32_t *m = static_cast<const int32_t *>(matrix);
            format_range(r0, r1, cols, names, name_off,
                         [m, ld](std::string &s, int64_t r, int32_t c) {
                             const int64_t v = m[r * ld + c];
                             if (v < 0) s.push_back('-');
                             put_uint(s, (uint64_t)(v < 0 ? -v : v));
                         },
                         &parts[t]);
        }
    };
    if (n_threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t) pool.emplace_back(work, t);
        for (auto &th : pool) th.join();
    }
    size_t total = 0;
    for (auto &p : parts) total += p.size();
    *written = total;
    if (total > cap || (!out && total))
        return sd::fail(SD_ERR_WORKSPACE, "sd_host_format_rows: %zu bytes needed, %zu given", total, cap);
    size_t off = 0;
    for (auto &p : parts) {
        memcpy(out + off, p.data(), p.size());
        off += p.size();
    }
    return SD_OK;
}
