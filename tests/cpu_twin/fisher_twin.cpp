// fisher_twin.cpp -- TEST-ONLY host build of splicedice_b200/csrc/sd_fisher_math.cuh.
// Lets the CPU test-suite check the kernel's per-table arithmetic against the scipy golden
// vectors without a GPU.  Not part of the product: nothing under splicedice_b200/ loads it.
#include <quadmath.h>
#include <stdint.h>

#include <vector>

#include "../../splicedice_b200/csrc/sd_fisher_math.cuh"
#include "../../splicedice_b200/csrc/sd_lgtable.h"

namespace {
struct HostTable {
    const double *t;
    int64_t n;
    double hi(int64_t k) const { return get(k).hi; }
    sd::fisher::dd in_table(int32_t k) const { return sd::fisher::dd_make(t[2 * k], t[2 * k + 1]); }
    sd::fisher::dd get(int64_t k) const
    {
        return k < n ? sd::fisher::dd_make(t[2 * k], t[2 * k + 1]) : sd::fisher::lgfact_stirling(*this, (double)k);
    }
};
}  // namespace

extern "C" int fisher_twin_batch(int64_t count, const int64_t *a, const int64_t *b, const int64_t *c,
                                 const int64_t *d, double *out, int64_t table_cap)
{
    int64_t nmax = 0;
    for (int64_t i = 0; i < count; ++i) {
        int64_t N = a[i] + b[i] + c[i] + d[i];
        if (N > nmax) nmax = N;
    }
    int64_t entries = nmax + 1;
    if (table_cap > 0 && entries > table_cap) entries = table_cap;
    std::vector<double> tab;
    sd::lgtable_host(entries, &tab);
    HostTable T{tab.data(), entries};
    for (int64_t i = 0; i < count; ++i) out[i] = (nmax < (int64_t(1) << 30) ? sd::fisher::two_sided<int32_t>(T, (int32_t)a[i], (int32_t)b[i], (int32_t)c[i], (int32_t)d[i]) : sd::fisher::two_sided<int64_t>(T, a[i], b[i], c[i], d[i]));
    return 0;
}

// log k! beyond the table (Stirling in double-double) against binary128 lgammaq: returns hi / lo words
// and the binary128 value split the same way
extern "C" void fisher_twin_stirling(int64_t count, const int64_t *k, double *hi, double *lo, double *want_hi,
                                     double *want_lo)
{
    std::vector<double> tab;
    sd::lgtable_host(4096, &tab);
    HostTable T{tab.data(), 4096};
    for (int64_t i = 0; i < count; ++i) {
        const sd::fisher::dd v = sd::fisher::lgfact_stirling(T, (double)k[i]);
        hi[i] = v.hi; lo[i] = v.lo;
        const __float128 w = lgammaq((__float128)k[i] + 1);
        want_hi[i] = (double)w;
        want_lo[i] = (double)(w - (__float128)want_hi[i]);
    }
}

// exp_small on an array (tests/test_fisher_twin.py checks it against binary128 / mpmath)
extern "C" void fisher_twin_exp(int64_t count, const double *x, double *out)
{
    for (int64_t i = 0; i < count; ++i) out[i] = sd::fisher::exp_small(x[i]);
}
