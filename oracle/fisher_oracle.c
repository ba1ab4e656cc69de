/* fisher_oracle.c -- binary128 restatement of scipy's two-sided Fisher exact test.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Built by
 * `__graft_entry__.build()` / `oracle/build_oracle.py` into oracle/_build/.
 *
 * The reference calls `scipy.stats.fisher_exact(table)[1]` per (event, sample pair)
 * (/root/reference/splicedice/pairwise_fisher.py:165,179).  scipy is a third-party
 * dependency that is not under /root/reference: the reference pins scipy==1.4.1
 * (requirements.txt:3), this image carries scipy 1.18.1, and the control flow
 * restated here is 1.18.1's `fisher_exact`, two-sided branch
 * (scipy/stats/_stats_py.py:5042-5108) with its binary search
 * (scipy/stats/_binomtest.py:342-386).  scipy evaluates pmf/cdf/sf with Boost.Math's
 * hypergeometric distribution in double precision; here every pmf is evaluated
 * directly from binary128 lgamma and every tail is summed in binary128 the way
 * Boost's `hypergeometric_cdf_imp` does (start at the boundary term, walk away
 * from the mode with the term ratio), so the result is scipy's answer with the
 * ~1e-15 Boost rounding noise removed.  tests/test_oracle_golden.py pins it to
 * scipy itself (and to mpmath exact sums) on the committed golden tables, and
 * tests/test_reference_live.py to scipy run live.
 */
#include <quadmath.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef __float128 q_t;

typedef struct {
    int64_t n1, n2, n, N;   /* scipy: n1 = c00+c01, n2 = c10+c11, n = c00+c10 (:5065-5067) */
    int64_t lo, hi;         /* support of hypergeom(M=N, n=n1, N=n) */
    const q_t *lg;          /* lg[k] = lgamma(k+1) = log k!  for k = 0..N */
} hyp_t;

static q_t *g_lg = NULL;
static int64_t g_lg_n = 0;

/* log k! table up to nmax (inclusive); grows monotonically, not thread-safe by design:
 * fisher_oracle_batch() extends it once before its parallel loop. */
static int ensure_table(int64_t nmax)
{
    if (nmax < g_lg_n) return 0;
    int64_t want = nmax + 1;
    q_t *t = (q_t *)realloc(g_lg, (size_t)want * sizeof(q_t));
    if (!t) return -1;
    for (int64_t k = g_lg_n; k < want; ++k) t[k] = lgammaq((q_t)k + 1.0Q);
    g_lg = t;
    g_lg_n = want;
    return 0;
}

static inline q_t log_choose(const q_t *lg, int64_t n, int64_t k)
{
    return lg[n] - lg[k] - lg[n - k];
}

/* hypergeom.pmf(x, M=N, n=n1, N=n): zero off the support */
static q_t pmf(const hyp_t *h, int64_t x)
{
    if (x < h->lo || x > h->hi) return 0.0Q;
    return expq(log_choose(h->lg, h->n1, x) + log_choose(h->lg, h->n2, h->n - x)
                - log_choose(h->lg, h->N, h->n));
}

/* lower tail  sum_{k<=x} pmf(k)  summed from x downwards (Boost hypergeometric_cdf_imp, x < mode) */
static q_t lower_tail(const hyp_t *h, int64_t x)
{
    if (x < h->lo) return 0.0Q;
    if (x > h->hi) x = h->hi;
    q_t term = pmf(h, x), sum = term;
    while (x > h->lo && term > sum * 1e-40Q) {
        /* pmf(x-1)/pmf(x) = x (n2 - n + x) / ((n1 - x + 1)(n - x + 1)) */
        term = term * (q_t)x * (q_t)(h->n2 - h->n + x) / ((q_t)(h->n1 - x + 1) * (q_t)(h->n - x + 1));
        sum += term;
        --x;
    }
    return sum;
}

/* upper tail  sum_{k>x} pmf(k)  summed from x+1 upwards (Boost, x >= mode branch) */
static q_t upper_tail(const hyp_t *h, int64_t x)
{
    if (x >= h->hi) return 0.0Q;
    if (x < h->lo) x = h->lo - 1;
    ++x;
    q_t term = pmf(h, x), sum = term;
    while (x < h->hi && term > sum * 1e-40Q) {
        /* pmf(x+1)/pmf(x) = (n1 - x)(n - x) / ((x + 1)(n2 - n + x + 1)) */
        term = term * (q_t)(h->n1 - x) * (q_t)(h->n - x) / ((q_t)(x + 1) * (q_t)(h->n2 - h->n + x + 1));
        sum += term;
        ++x;
    }
    return sum;
}

/* scipy _binary_search_for_binom_tst (_binomtest.py:342-386), sign = +1 searches a(x)=pmf(x),
 * sign = -1 searches a(x) = -pmf(x). */
static int64_t binary_search(const hyp_t *h, int sign, q_t d, int64_t lo, int64_t hi)
{
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        q_t v = (q_t)sign * pmf(h, mid);
        if (v < d) lo = mid + 1;
        else if (v > d) hi = mid - 1;
        else { lo = mid; hi = mid; }
    }
    return ((q_t)sign * pmf(h, lo) <= d) ? lo : lo - 1;
}

static double fisher_two_sided(int64_t c00, int64_t c01, int64_t c10, int64_t c11)
{
    hyp_t h;
    h.n1 = c00 + c01; h.n2 = c10 + c11; h.n = c00 + c10; h.N = h.n1 + h.n2;
    /* a zero row or column sum: p = 1 (_stats_py.py:5055-5058) */
    if (h.n1 == 0 || h.n2 == 0 || h.n == 0 || c01 + c11 == 0) return 1.0;
    h.lo = h.n - h.n2 > 0 ? h.n - h.n2 : 0;
    h.hi = h.n1 < h.n ? h.n1 : h.n;
    h.lg = g_lg;

    /* mode = int((n + 1) * (n1 + 1) / (n1 + n2 + 2))   (:5078; float divide then truncation) */
    double modef = (double)((h.n + 1) * (h.n1 + 1)) / (double)(h.N + 2);
    int64_t mode = (int64_t)modef;
    q_t pexact = pmf(&h, c00), pmode = pmf(&h, mode);
    const q_t eps = 1e-14Q;
    const q_t gamma = (q_t)(1.0 + 1e-14);           /* scipy forms 1 + epsilon in double (:5083) */
    q_t big = pexact > pmode ? pexact : pmode;
    if (fabsq(pexact - pmode) / big <= eps) return 1.0;           /* :5085-5086 */

    q_t p;
    if (c00 < mode) {                                              /* :5088-5094 */
        q_t plower = lower_tail(&h, c00);
        if (pmf(&h, h.n) > pexact * gamma) p = plower;
        else {
            int64_t guess = binary_search(&h, -1, -pexact * gamma, mode, h.n);
            p = plower + upper_tail(&h, guess);
        }
    } else {                                                       /* :5095-5101 */
        q_t pupper = upper_tail(&h, c00 - 1);
        if (pmf(&h, 0) > pexact * gamma) p = pupper;
        else {
            int64_t guess = binary_search(&h, +1, pexact * gamma, 0, mode);
            p = pupper + lower_tail(&h, guess);
        }
    }
    if (p > 1.0Q) p = 1.0Q;                                        /* :5106 */
    return (double)p;
}

/* out[i] = two-sided p of [[a[i], b[i]], [c[i], d[i]]].  Returns 0, or -1 on allocation
 * failure / negative input (scipy raises ValueError on negatives, :5048-5049). */
int fisher_oracle_batch(int64_t count, const int64_t *a, const int64_t *b,
                        const int64_t *c, const int64_t *d, double *out)
{
    int64_t nmax = 0;
    for (int64_t i = 0; i < count; ++i) {
        if (a[i] < 0 || b[i] < 0 || c[i] < 0 || d[i] < 0) return -1;
        int64_t N = a[i] + b[i] + c[i] + d[i];
        if (N > nmax) nmax = N;
    }
    if (ensure_table(nmax + 2) != 0) return -1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < count; ++i)
        out[i] = fisher_two_sided(a[i], b[i], c[i], d[i]);
    return 0;
}

/* hypergeometric support size K = hi - lo + 1 (0 for trivial tables): the work unit of the
 * FP64 roofline model. */
int fisher_oracle_support(int64_t count, const int64_t *a, const int64_t *b,
                          const int64_t *c, const int64_t *d, int64_t *k_out)
{
    for (int64_t i = 0; i < count; ++i) {
        int64_t n1 = a[i] + b[i], n2 = c[i] + d[i], n = a[i] + c[i];
        if (n1 == 0 || n2 == 0 || n == 0 || b[i] + d[i] == 0) { k_out[i] = 0; continue; }
        int64_t lo = n - n2 > 0 ? n - n2 : 0, hi = n1 < n ? n1 : n;
        k_out[i] = hi - lo + 1;
    }
    return 0;
}
