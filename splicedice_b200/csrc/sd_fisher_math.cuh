// sd_fisher_math.cuh -- per-table arithmetic of the two-sided Fisher exact test.
//
// Follows the control flow of scipy.stats.fisher_exact (scipy 1.18.1,
// stats/_stats_py.py:5042-5108; the call the reference makes at
// /root/reference/splicedice/pairwise_fisher.py:165,179):
//   - any zero margin                         -> p = 1                         (:5055-5058)
//   - mode = int((n+1)(n1+1)/(n1+n2+2))       (float divide, truncation)       (:5078)
//   - |pexact - pmode| / max <= 1e-14         -> p = 1                         (:5085-5086)
//   - observed below the mode: p = cdf(a) + sum of the upper-side x with pmf(x) < pexact*(1+1e-14)
//     observed above the mode: p = sf(a-1) + sum of the lower-side x with pmf(x) <= pexact*(1+1e-14)
//                                                                              (:5088-5101)
//   - p = min(p, 1)                                                            (:5106)
// The pmf itself is never evaluated per support point.  With t(x) = pmf(x)/pmf(a):
//   p = pmf(a) * ( S(a, away from the mode) + t(g) * S(g, away from the mode) )
// where g is the first far-side point that scipy's search admits and S(x0, dir) =
// 1 + r1 + r1 r2 + ... is the tail sum relative to its first term, accumulated with the exact
// rational term ratio  pmf(x+1)/pmf(x) = (n1-x)(n-x) / ((x+1)(n2-n+x+1))  as a
// numerator/denominator pair (no division in the loop) and cut when a term drops below
// 2^-kCutBits of the running sum.  pmf(a), t(g) and every "is pmf(x) within the tie window of
// pmf(a)" decision come from a double-double log-factorial table (sd_lgtable.cpp).
//
// The header compiles for the device (sd_fisher.cu) and, without nvcc, for the host: the host
// build exists only so tests/ can check this arithmetic against scipy golden vectors on
// machines with no GPU.  Nothing in the product calls the host build.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SD_HD __host__ __device__ __forceinline__
#define SD_NOINLINE __host__ __device__ __noinline__
#else
#define SD_HD inline
#define SD_NOINLINE inline
#endif

namespace sd {
namespace fisher {

struct dd {
    double hi, lo;
};

SD_HD dd dd_make(double hi, double lo) { dd r; r.hi = hi; r.lo = lo; return r; }

SD_HD dd dd_add(dd x, dd y)
{
    double s = x.hi + y.hi;
    double bb = s - x.hi;
    double e = (x.hi - (s - bb)) + (y.hi - bb);
    e += x.lo + y.lo;
    double hi = s + e;
    return dd_make(hi, e - (hi - s));
}
SD_HD dd dd_sub(dd x, dd y) { return dd_add(x, dd_make(-y.hi, -y.lo)); }

// (sum, err) += (h, l): error-free TwoSum of the leading words, everything else into err
SD_HD void acc_two_sum(double &sum, double &err, double h, double l)
{
    const double s = sum + h;
    const double bb = s - sum;
    err += ((sum - (s - bb)) + (h - bb)) + l;
    sum = s;
}

constexpr double kEps = 1e-14;                          // scipy's epsilon (:5082)
constexpr double kLogGamma = 9.992007221626358e-15;     // log(1 + 1e-14) with 1 + 1e-14 formed in binary64
constexpr double kTieWindow = 1.0000000000000051e-14;   // -log(1 - 1e-14)
// Tail truncation (tables whose total is below 2^26, the second-difference form): a tail stops
// once its term falls below 2^-kCutBits of the running sum and the rest of the tail is added in
// closed form -- the geometric series of the next term ratio r, term * r / (1 - r).  The true
// ratios keep falling, so that overestimates the remainder by about 1 / z^2 of it (z = distance
// from the mode in standard deviations, ~7 where a 2^-38 term sits): with the remainder itself
// at 2^-38 * sigma / z of the sum and sigma <= 2^11 for these totals, the error left stays below
// 2e-11 of the sum in the worst case and is far smaller on real counts.  Measured against the
// binary128 oracle (CPU twin, random tables with cells up to 60 / 700 / 9,000 / 150,000 / 3e6 and
// the configs[2] distribution, maximum relative error):
//     48 bits  8e-16 .. 3e-15        40 bits  1e-15 .. 4e-13        36 bits  2e-14 .. 8e-12
//     32 bits  5e-13 .. 2e-10        28 bits  2e-11 .. 3e-9
// and on the B200 (200,000 x 2,016, ms per launch): 48 / 40 / 38 / 36 bits = 20.33 / 19.28 /
// 19.03 / 18.77.  38 bits keeps a 5x margin under the 1e-11 the parity tests hold (the contract
// is 1e-9).  Round 1 cut at 2^-48 with no closed form: 20.0 ms.  The general form (tail_sum,
// totals from 2^26) keeps 2^-48 and no closed form.
#ifndef SD_FISHER_CUT_BITS
#define SD_FISHER_CUT_BITS 38
#endif
constexpr int kCutBits = SD_FISHER_CUT_BITS;
constexpr double kCut = 3.552713678800501e-15;          // 2^-48: tail_sum's truncation relative to the sum
constexpr double kBig = 3.273390607896142e150;          // 2^500
constexpr double kSmall = 3.054936363499605e-151;       // 2^-500

// Tail sum relative to its first term.  (p, q) are the two table cells that shrink along the
// walk, (u, v) the two that grow:  term_k / term_{k-1} = (p-k+1)(q-k+1) / ((u+k)(v+k)).
SD_HD double tail_sum(double p, double q, double u, double v)
{
    double steps = p < q ? p : q;                  // the ratio is zero after min(p, q) steps
    double num_p = p, num_q = q, den_u = u + 1.0, den_v = v + 1.0;
    double P = 1.0, Q = 1.0, A = 1.0;             // term = P / Q, sum = A / Q
    while (steps > 0.0) {
        int burst = steps < 4.0 ? (int)steps : 4;
        for (int i = 0; i < burst; ++i) {
            P *= num_p * num_q;
            double den = den_u * den_v;
            Q *= den;
            A = fma(A, den, P);
            num_p -= 1.0; num_q -= 1.0; den_u += 1.0; den_v += 1.0;
        }
        steps -= (double)burst;
        if (P < kCut * A) break;
        if (Q > kBig) { P *= kSmall; Q *= kSmall; A *= kSmall; }
    }
    return A / Q;
}

// Tail sum for tables whose total is below 2^26 (every product below stays an exact integer
// under 2^52).  The numerator (p-k)(q-k) and denominator (u+1+k)(v+1+k) of the term ratio are
// quadratics in k, advanced by second differences: two additions each instead of two additions
// and a multiplication, 7 FP64 operations per term.  A tail that runs out of support multiplies
// its term by zero and stops at the next check.
// high 32 bits of a binary64 (sign, exponent, top 20 fraction bits): for positive values an integer
// comparison of these words orders the values to within 2^-20 relative
SD_HD int32_t hi_word(double x)
{
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    int64_t bits;
    memcpy(&bits, &x, sizeof bits);
    return (int32_t)(bits >> 32);
#endif
}
// x * 2^-500 for x >= 2^500 (exact): an integer subtraction on the exponent field on the device
SD_HD double scale_down(double x)
{
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(x) - (500 << 20), __double2loint(x));
#else
    return x * kSmall;
#endif
}
constexpr int32_t kBigHi = (1023 + 500) << 20;      // hi_word(2^500)
constexpr int32_t kCutHi = kCutBits << 20;          // 2^-kCutBits in hi_word units

struct TailState {
    double num, s, den, w, P, Q, A;
    SD_HD void init(double p, double q, double u, double v)
    {
        num = p * q; s = p + q - 1.0;               // (p-k-1)(q-k-1) = num_k - s_k,  s_{k+1} = s_k - 2
        den = (u + 1.0) * (v + 1.0); w = u + v + 3.0;   // (u+k+2)(v+k+2) = den_k + w_k,  w_{k+1} = w_k + 2
        P = 1.0; Q = 1.0; A = 1.0;
    }
    SD_HD void init_empty() { num = 0.0; s = 0.0; den = 1.0; w = 0.0; P = 0.0; Q = 1.0; A = 0.0; }
    SD_HD void step()
    {
        P *= num; Q *= den; A = fma(A, den, P);
        num -= s; s -= 2.0; den += w; w += 2.0;
    }
    // The cut and the rescale test compare exponent words with integer instructions (the FP64
    // pipe is the bottleneck): the tail stops once the term is below 2^-kCutBits (to within the 20
    // fraction bits of the word) of the running sum; P >= 0, A >= Q >= 1 always.
    SD_HD bool done() const { return hi_word(P) < hi_word(A) - kCutHi; }
    // sum / first term, with the terms not summed replaced by the geometric series of the next
    // ratio num / den (binary32 is plenty: the correction is below 2^-30 of the sum).  A tail that
    // ran out of support has P = 0.
    SD_HD double result() const
    {
        const float fn = (float)num, fd = (float)(den - num);
#ifdef __CUDA_ARCH__
        const float r = __fdividef(fn, fd);
#else
        const float r = fn / fd;
#endif
        return fma(P, (double)r, A) / Q;
    }
    SD_HD void rescale()
    {
        if (hi_word(Q) >= kBigHi) { P *= kSmall; Q = scale_down(Q); A = scale_down(A); }
    }
};

// one tail, relative to its first term
SD_HD double tail_fast(double p, double q, double u, double v)
{
    TailState t;
    t.init(p, q, u, v);
    // The cut is tested every four terms, the rescale only every eight: a step multiplies Q by
    // less than 2^52 (totals below 2^26), so from below 2^500 eight steps stay under 2^916.
    do {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < 4; ++i) t.step();
        if (t.done()) break;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < 4; ++i) t.step();
        t.rescale();
    } while (!t.done());
    return t.result();
}

// exp(x) for x <= ~1 (log-probabilities): n = rint(32 x / ln 2), r = x - n ln 2 / 32 in two FMA
// steps (|r| <= ln 2 / 64 = 0.0108), exp(x) = 2^(n >> 5) * T[n & 31] * (1 + p(r)) with T[j] =
// 2^(j/32) and p the degree-6 Taylor polynomial of expm1 (truncation r^7/5040 < 4e-18), scaling
// by the power of two on the exponent field.  ~1 ulp, 12 FP64 operations (a degree-13 polynomial
// on |r| <= 0.347 needed 19).  x < -746 gives 0, results below 2^-1022 are rounded once by a
// final multiply.  On the device T sits in global memory behind the read-only cache (256 bytes,
// per-lane index: the constant bank would serialise the lanes); the polynomial coefficients are
// immediates / constant-bank operands of the DFMAs.
#ifdef __CUDA_ARCH__
#define SD_EXP_TAB __device__ const
#else
#define SD_EXP_TAB static const
#endif
SD_EXP_TAB double kExp2Tab[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,
    1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,
    1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.645755478153965,
    1.681792830507429, 1.718619298122478, 1.7562521603732995, 1.7947090750031072,
    1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002};

SD_HD double add_exponent(double x, int k)      // x * 2^k, result normal
{
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(x) + (k << 20), __double2loint(x));
#else
    return ldexp(x, k);
#endif
}

SD_HD double exp_small(double x)
{
    if (!(x >= -746.0)) return 0.0;
    const double kMagic = 6755399441055744.0;                  // 2^52 + 2^51: rint() in the low fraction bits
    const double t = fma(x, 46.16624130844683, kMagic);        // 32 / ln 2
    const double nf = t - kMagic;
    double r = fma(nf, -0.02166084939249829, x);               // ln 2 / 32, hi and lo
    r = fma(nf, -7.247021293269686e-19, r);
    double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p *= r;                                                    // expm1(r)
    const int n = (int)nf;
#ifdef __CUDA_ARCH__
    const double tj = __ldg(&kExp2Tab[n & 31]);
#else
    const double tj = kExp2Tab[n & 31];
#endif
    const double q = fma(tj, p, tj);
    const int k = n >> 5;                                      // floor(n / 32): arithmetic shift
    if (k >= -1020) return add_exponent(q, k);
    return add_exponent(q, k + 1000) * 9.3326361850321888e-302;   // 2^-1000
}

// (hi, lo) of a compensated sum: renormalise the TwoSum accumulator pair
SD_HD dd acc_result(double sum, double err)
{
    const double hi = sum + err;
    return dd_make(hi, err - (hi - sum));
}

// log k! for k beyond the log-factorial table, in double-double: Stirling's series
//   log k! = (k + 1/2) log k - k + log(2 pi)/2 + 1/(12 k) - 1/(360 k^3) + ...
// (k >= 4096 here: the k^-3 term is below 7e-14 and kept, the k^-5 term below 7e-22 is dropped;
// the table itself covers 2^22 entries, where both vanish).  A binary64 lgamma() is not enough: at
// k = 4e6, log k! = 5.7e7 carries 7e-9 of rounding error, more than the 1e-9 relative the p-values
// are promised to.  log k comes from the table itself: k = 2^s (c + delta) with c an integer in
// [1024, 2048] and |delta| <= 1/2, so
//   log k = s ln 2 + (lg[c] - lg[c-1]) + log1p(delta / c),    |delta / c| <= 2^-11,
// the quotient and its square carried in double-double, the series to the eighth power.  The
// result is good to ~1e-14 absolute at k = 2^33 (1e-17 at 2^23); checked against binary128
// lgammaq in tests/test_fisher_twin.py.  Out of line: this is the rare path.
template <class Table>
SD_NOINLINE dd lgfact_stirling(Table tab, double k)
{
    int e;
    const double m = frexp(k, &e) * 2048.0;                  // [1024, 2048), exact
    const double c = rint(m);
    const double delta = m - c;                              // exact (Sterbenz)
    const int s = e - 11;
    // r = delta / c in double-double
    const double r_hi = delta / c;
    const double r_lo = fma(-r_hi, c, delta) / c;
    // log1p(r) = r - r^2/2 + r^3/3 - ... - r^8/8
    const double sq_hi = r_hi * r_hi;
    const double sq_lo = fma(r_hi, r_hi, -sq_hi) + 2.0 * r_hi * r_lo;
    double t = fma(r_hi, -1.0 / 8.0, 1.0 / 7.0);
    t = fma(t, r_hi, -1.0 / 6.0);
    t = fma(t, r_hi, 1.0 / 5.0);
    t = fma(t, r_hi, -1.0 / 4.0);
    t = fma(t, r_hi, 1.0 / 3.0);
    t *= sq_hi * r_hi;                                       // r^3/3 - ... : below 4e-11, plain double
    double sum = r_hi, err = r_lo + t;
    acc_two_sum(sum, err, -0.5 * sq_hi, -0.5 * sq_lo);
    // + log c = lg[c] - lg[c-1]
    const dd lc = tab.in_table((int32_t)c), lc1 = tab.in_table((int32_t)c - 1);   // every table holds >= 4096 entries
    acc_two_sum(sum, err, lc.hi, lc.lo);
    acc_two_sum(sum, err, -lc1.hi, -lc1.lo);
    // + s ln 2 (s < 64: the product with the hi word is exact to 6 spare bits only if split --
    // use an FMA pair instead)
    {
        const double sd = (double)s;
        const double p_hi = sd * 0.6931471805599453;
        const double p_lo = fma(sd, 0.6931471805599453, -p_hi) + sd * 2.3190468138462996e-17;
        acc_two_sum(sum, err, p_hi, p_lo);
    }
    const dd logk = acc_result(sum, err);
    // (k + 1/2) * log k: k + 1/2 is exact below 2^52
    const double kh = k + 0.5;
    const double q_hi = kh * logk.hi;
    const double q_lo = fma(kh, logk.hi, -q_hi) + kh * logk.lo;
    sum = q_hi; err = q_lo;
    acc_two_sum(sum, err, -k, 0.0);
    acc_two_sum(sum, err, 0.9189385332046728, -3.8782941580672414e-17);
    const double ik = 1.0 / k;
    acc_two_sum(sum, err, ik * (1.0 / 12.0), -ik * ik * ik * (1.0 / 360.0));
    return acc_result(sum, err);
}

// Per-sample terms of a table column (inc, exc): both depend on one sample only, so the pairwise
// kernel computes them once per (junction, sample) instead of once per pair.  Kept as unnormalised
// (sum, err) pairs of a compensated sum -- TwoSum on the hi words, rounding errors and lo words
// accumulated in err:
//   g = lg[inc] + lg[exc]                 (this column's half of G(a))
//   c = lg[inc + exc] - lg[inc] - lg[exc] (log of the binomial coefficient C(inc + exc, inc))
struct SampleTerms {
    double g_sum, g_err, c_sum, c_err;
};

template <class Table, class Int>
SD_HD SampleTerms sample_terms(const Table &tab, Int inc, Int exc)
{
    const dd ti = tab.get(inc), te = tab.get(exc), tn = tab.get(inc + exc);
    SampleTerms t;
    t.g_sum = ti.hi; t.g_err = ti.lo;
    acc_two_sum(t.g_sum, t.g_err, te.hi, te.lo);
    t.c_sum = tn.hi; t.c_err = tn.lo;
    acc_two_sum(t.c_sum, t.c_err, -ti.hi, -ti.lo);
    acc_two_sum(t.c_sum, t.c_err, -te.hi, -te.lo);
    return t;
}

// f(x) = log(pmf(x) / pmf(a)) in double-double: G(a) - G(x) as one compensated sum -- G(a) comes
// in as its (sum, err) pair, the four table entries of G(x) are subtracted with TwoSum on the hi
// words, rounding errors and lo words accumulated separately (the error words are ~1e-12, so
// their plain-double sum is good to ~1e-27).  Out of line, with everything passed by value, so
// the call does not force the caller's state into local memory.
template <class Table, class Int>
SD_NOINLINE dd f_exact_of(Table tab, double ga_sum, double ga_err, Int n1, Int n2, Int n, Int x)
{
    const dd t1 = tab.get(x), t3 = tab.get(n1 - x), t5 = tab.get(n - x), t7 = tab.get(n2 - n + x);
    double sum = ga_sum, err = ga_err;
    acc_two_sum(sum, err, -t1.hi, -t1.lo);
    acc_two_sum(sum, err, -t3.hi, -t3.lo);
    acc_two_sum(sum, err, -t5.hi, -t5.lo);
    acc_two_sum(sum, err, -t7.hi, -t7.lo);
    return acc_result(sum, err);
}

// G(x) = lg[x] + lg[n1-x] + lg[n-x] + lg[n2-n+x]; log pmf(x) = const - G(x).
// Int is the integer type of the table entries (int32_t when every table total fits 31 bits).
template <class Table, class Int>
struct Problem {
    const Table &tab;
    Int n1, n2, n;
    double ga_sum, ga_err;   // G(a) = lg[a] + lg[b] + lg[c] + lg[d] as a compensated (sum, err) pair
    double tol;   // bound on the error of a hi-only evaluation of G(a) - G(x)

    // f(x) = log(pmf(x) / pmf(a)) from the hi words only: at most seven roundings of at most
    // eps/2 * lg[N] each (every partial sum is below lg[N]), well inside `tol`
    SD_HD double f_fast(Int x) const
    {
        return ga_sum - ((tab.hi(x) + tab.hi(n1 - x)) + (tab.hi(n - x) + tab.hi(n2 - n + x)));
    }
    SD_HD dd f_exact(Int x) const { return f_exact_of<Table, Int>(tab, ga_sum, ga_err, n1, n2, n, x); }
    // scipy's far-side admission test: pmf(x) <= pexact * (1 + 1e-14)
    SD_HD bool admitted(Int x) const
    {
        const double f = f_fast(x);
        if (fabs(f - kLogGamma) > tol) return f < kLogGamma;
        const dd fe = f_exact(x);
        return fe.hi + fe.lo <= kLogGamma;
    }
};

// Everything about a table except its two tail sums.  `known`: the p-value is already decided
// (zero margin, observed == mode, tie with the mode) and sits in `pexact`.  Otherwise
//   p = pexact * (S(near tail) + tg * S(far tail)),
// each tail described by its four starting cells (p, q shrink; u, v grow); `far[0] < 0` when the
// far side contributes nothing.
template <class Int>
struct Plan {
    bool known;
    double pexact, tg;
    Int near[4], far[4];
    Int total;              // N, decides which tail-sum routine is exact
};

// ta / tb: sample_terms of the table's two columns, (a, c) and (b, d)
template <class Int, class Table>
SD_HD Plan<Int> make_plan_pre(const Table &tab, Int a, Int b, Int c, Int d, const SampleTerms &ta, const SampleTerms &tb)
{
    Plan<Int> pl;
    pl.known = true; pl.pexact = 1.0; pl.tg = 0.0; pl.total = 0;
    pl.near[0] = pl.near[1] = pl.near[2] = pl.near[3] = 0;
    pl.far[0] = -1; pl.far[1] = pl.far[2] = pl.far[3] = 0;
    Int n1 = a + b, n2 = c + d, n = a + c;
    if (n1 == 0 || n2 == 0 || n == 0 || b + d == 0) return pl;
    const Int N = n1 + n2;
    // numpy: int64 product, float64 divide, int() truncation
    Int mode = (Int)((double)((int64_t)(n + 1) * (int64_t)(n1 + 1)) / (double)((int64_t)N + 2));
    if (a == mode) return pl;
    // Observed count above the mode: swap the columns.  x -> n1 - x maps the distribution onto
    // hypergeom(N, n1, N - n) with identical pmf values, so from here on a < mode and the far
    // side is the upper one.  (The reflected mode is only used as a point whose pmf exceeds
    // pmf(a) * (1 + 1e-14), which the tie test below guarantees.)  Everything taken from ta / tb
    // is symmetric in the two columns.
    if (a > mode) {
        Int t = a; a = b; b = t;
        t = c; c = d; d = t;
        n = N - n;
        mode = n1 - mode;
    }
    const Int hi = n1 < n ? n1 : n;

    Problem<Table, Int> pr{tab, n1, n2, n, ta.g_sum, ta.g_err, 0.0};
    acc_two_sum(pr.ga_sum, pr.ga_err, tb.g_sum, tb.g_err);
    pr.tol = 64.0 * 2.220446049250313e-16 * (tab.hi(N) + 1.0);

    // tie between the observed table and the mode (plateau or mirror twin)
    {
        const double fm = pr.f_fast(mode);
        if (fabs(fm) <= pr.tol + kTieWindow) {
            const dd fe = pr.f_exact(mode);
            if (fabs(fe.hi + fe.lo) <= kTieWindow) return pl;
        }
    }

    // log pmf(a) = lg[n1] + lg[n2] - lg[N] + log C(a + c, a) + log C(b + d, b): compensated sum
    // of three table entries and the two per-sample terms
    double lp_hi, lp_lo;
    {
        const dd t0 = tab.get(n1), t1 = tab.get(n2), t4 = tab.get(N);
        double sum = t0.hi, err = t0.lo;
        acc_two_sum(sum, err, t1.hi, t1.lo);
        acc_two_sum(sum, err, -t4.hi, -t4.lo);
        acc_two_sum(sum, err, ta.c_sum, ta.c_err);
        acc_two_sum(sum, err, tb.c_sum, tb.c_err);
        lp_hi = sum + err;
        lp_lo = err - (lp_hi - sum);
    }

    // smallest g in (mode, hi] with admitted(g): start from the mirror image of a, gallop, bisect
    Int lo_x = mode, hi_x = hi + 1;
    {
        Int x = 2 * mode - a;
        if (x > hi) x = hi;
        if (x <= mode) x = mode + 1;
        if (x <= hi) {
            const bool up = !pr.admitted(x);        // true: the boundary is above x
            if (up) lo_x = x; else hi_x = x;
            Int step = 1;
            while (true) {
                const Int y = up ? lo_x + step : hi_x - step;
                if (up ? y >= hi_x : y <= lo_x) break;
                const bool adm = pr.admitted(y);
                if (adm) hi_x = y; else lo_x = y;
                if (adm == up) break;               // the boundary is bracketed
                step *= 2;
            }
            while (hi_x - lo_x > 1) {
                const Int mid = lo_x + (hi_x - lo_x) / 2;
                if (pr.admitted(mid)) hi_x = mid; else lo_x = mid;
            }
        }
    }
    const Int g = hi_x;
    pl.known = false;
    pl.total = N;
    pl.near[0] = a; pl.near[1] = d; pl.near[2] = b; pl.near[3] = c;                  // a, a-1, ...
    if (g <= hi) {
        const dd fg = pr.f_exact(g);
        double tg = exp_small(fg.hi);
        pl.tg = fma(tg, fg.lo, tg);
        pl.far[0] = n1 - g; pl.far[1] = n - g; pl.far[2] = g; pl.far[3] = n2 - n + g;
    }
    double pexact = exp_small(lp_hi);
    pl.pexact = fma(pexact, lp_lo, pexact);
    return pl;
}

template <class Int, class Table>
SD_HD Plan<Int> make_plan(const Table &tab, Int a, Int b, Int c, Int d)
{
    return make_plan_pre<Int>(tab, a, b, c, d, sample_terms(tab, a, c), sample_terms(tab, b, d));
}

template <class Int>
SD_HD double plan_value(const Plan<Int> &pl)
{
    if (pl.known) return pl.pexact;
    double rel;
    if ((int64_t)pl.total < (int64_t(1) << 26)) {
        rel = tail_fast((double)pl.near[0], (double)pl.near[1], (double)pl.near[2], (double)pl.near[3]);
        if (pl.far[0] >= 0)
            rel = fma(pl.tg, tail_fast((double)pl.far[0], (double)pl.far[1], (double)pl.far[2], (double)pl.far[3]), rel);
    } else {
        rel = tail_sum((double)pl.near[0], (double)pl.near[1], (double)pl.near[2], (double)pl.near[3]);
        if (pl.far[0] >= 0)
            rel = fma(pl.tg, tail_sum((double)pl.far[0], (double)pl.far[1], (double)pl.far[2], (double)pl.far[3]), rel);
    }
    const double p = pl.pexact * rel;
    return p > 1.0 ? 1.0 : p;
}

template <class Int, class Table>
SD_HD double two_sided(const Table &tab, Int a, Int b, Int c, Int d)
{
    return plan_value(make_plan<Int>(tab, a, b, c, d));
}

// the pairwise kernel's form: the columns' per-sample terms were computed once per junction
template <class Int, class Table>
SD_HD double two_sided_pre(const Table &tab, Int a, Int b, Int c, Int d, const SampleTerms &ta, const SampleTerms &tb)
{
    return plan_value(make_plan_pre<Int>(tab, a, b, c, d, ta, tb));
}

// hypergeometric support size of a table (0 for a zero-margin table): the work unit of the
// FP64 model in DESIGN.md.
SD_HD int64_t support_size(int64_t a, int64_t b, int64_t c, int64_t d)
{
    const int64_t n1 = a + b, n2 = c + d, n = a + c;
    if (n1 == 0 || n2 == 0 || n == 0 || b + d == 0) return 0;
    const int64_t lo = n - n2 > 0 ? n - n2 : 0, hi = n1 < n ? n1 : n;
    return hi - lo + 1;
}

}  // namespace fisher
}  // namespace sd
