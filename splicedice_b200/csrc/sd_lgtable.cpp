// sd_lgtable.cpp -- host side of the log-factorial table used by the Fisher kernel.
//
// lg[k] = log(k!) as an unevaluated sum hi + lo of two binary64 numbers (~106 bits), computed in
// binary128 on the host (libquadmath) by a running sum of logq(k).  Two-sided Fisher p-values
// hinge on comparing pmf(x) with pmf(observed) (scipy: stats/_stats_py.py:5082-5101, relative
// window 1e-14); a plain binary64 table of log k! carries ~1e-12 absolute error at k ~ 4000,
// which is enough to put an exact mirror tie on the wrong side.  The double-double table makes
// those comparisons exact for every table total below the table size.
#include <quadmath.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "sd_lgtable.h"

namespace sd {

namespace {
std::mutex g_mu;
std::vector<double> g_host;          // interleaved hi, lo
__float128 g_last = 0;               // log((n-1)!) for n = g_host.size() / 2
}  // namespace

// Extends the process-wide host table to at least `entries` entries and returns a pointer to
// it (interleaved hi/lo).  The returned storage is only ever appended to under the lock; the
// caller copies out of it while holding no reference across calls.
void lgtable_host(int64_t entries, std::vector<double> *out)
{
    std::lock_guard<std::mutex> lock(g_mu);
    int64_t have = (int64_t)g_host.size() / 2;
    if (have < entries) {
        g_host.reserve((size_t)entries * 2);
        for (int64_t k = have; k < entries; ++k) {
            if (k >= 2) g_last += logq((__float128)k);
            double hi = (double)g_last;
            double lo = (double)(g_last - (__float128)hi);
            g_host.push_back(hi);
            g_host.push_back(lo);
        }
    }
    out->assign(g_host.begin(), g_host.begin() + (size_t)entries * 2);
}

}  // namespace sd
