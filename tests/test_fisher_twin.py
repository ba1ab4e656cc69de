"""The kernel's per-table Fisher arithmetic (splicedice_b200/csrc/sd_fisher_math.cuh) compiled
for the host -- TEST-ONLY build under tests/cpu_twin -- against the scipy golden tables, the
exact rational sums and the binary128 oracle.  Lets the CPU suite catch arithmetic regressions
without a GPU; the GPU suite runs the same vectors through the real kernel."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import fisher_c
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def twin():
    out_dir = os.path.join(HERE, "cpu_twin", "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libfisher_twin.so")
    srcs = [os.path.join(HERE, "cpu_twin", "fisher_twin.cpp"),
            os.path.join(ROOT, "splicedice_b200", "csrc", "sd_lgtable.cpp")]
    deps = srcs + [os.path.join(ROOT, "splicedice_b200", "csrc", "sd_fisher_math.cuh")]
    if not os.path.isfile(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-o", lib, *srcs,
                        "-lquadmath", "-lm"], check=True)
    handle = ctypes.CDLL(lib)
    i64p = ctypes.POINTER(ctypes.c_int64)

    def run(tables, cap=0):
        t = np.ascontiguousarray(tables, dtype=np.int64)
        cols = [np.ascontiguousarray(t[:, i]) for i in range(4)]
        out = np.empty(len(t))
        handle.fisher_twin_batch(ctypes.c_int64(len(t)), *[c.ctypes.data_as(i64p) for c in cols],
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_int64(cap))
        return out

    def exp_small(x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty(len(x))
        handle.fisher_twin_exp(ctypes.c_int64(len(x)), x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                               out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        return out
    def stirling(k):
        k = np.ascontiguousarray(k, dtype=np.int64)
        outs = [np.empty(len(k)) for _ in range(4)]
        dp = ctypes.POINTER(ctypes.c_double)
        handle.fisher_twin_stirling(ctypes.c_int64(len(k)), k.ctypes.data_as(i64p), *[o.ctypes.data_as(dp) for o in outs])
        return outs
    run.exp_small = exp_small
    run.stirling = stirling
    return run


def _max_rel(got, want):
    ok = want > 1e-300
    return float((np.abs(got[ok] - want[ok]) / want[ok]).max(initial=0.0))


def test_scipy_golden(twin):
    g = util.load_npz("fisher_tables.npz")
    got = twin(g["tables"])
    assert _max_rel(got, g["p"]) < 1e-9
    assert np.array_equal(got[g["p"] == 1.0], g["p"][g["p"] == 1.0])


def test_exact_rational(twin):
    g = util.load_npz("fisher_exact_small.npz")
    np.testing.assert_allclose(twin(g["tables"]), g["p"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("scale,n", [(5, 50000), (60, 50000), (700, 30000), (9000, 8000), (150000, 400)])
def test_random_vs_binary128(twin, scale, n):
    rng = np.random.default_rng(scale)
    t = rng.integers(0, scale, size=(n, 4))
    t[::4, 0] *= 3
    want = fisher_c.fisher_two_sided(t[:, 0], t[:, 1], t[:, 2], t[:, 3])
    assert _max_rel(twin(t), want) < 1e-11


def test_log_factorial_beyond_the_table_is_double_double(twin):
    """fisher::lgfact_stirling (what the kernel uses for k >= 2^22, the table cap) against binary128
    lgammaq: the hi word is the correctly rounded double and hi + lo is within 2e-14 absolute up to
    2^33 (3e-16 relative to nothing: log k! is ~1e11 there) -- a plain lgamma() is 7e-9 off at 4e6,
    more than the 1e-9 the p-values are promised to."""
    rng = np.random.default_rng(1)
    k = np.unique(np.concatenate([
        np.arange(4096, 4200), 2 ** np.arange(12, 34), 2 ** np.arange(12, 34) - 1, 2 ** np.arange(12, 34) + 1,
        rng.integers(4096, 1 << 22, 3000), rng.integers(1 << 22, 1 << 26, 3000), rng.integers(1 << 26, 1 << 33, 3000)]))
    hi, lo, want_hi, want_lo = twin.stirling(k)
    err = np.abs((hi - want_hi) + (lo - want_lo))
    assert (hi == want_hi).all()
    assert err[k < (1 << 26)].max() < 2e-16 and err.max() < 2e-14
    plain = np.abs(np.array([float(__import__("math").lgamma(x + 1.0)) for x in k[-200:]]) - want_hi[-200:])
    assert plain.max() > 1e-7                                   # what the double-double form is for


def test_beyond_the_table_cap(twin):
    """Totals above the table (cap forced to 4,096 entries, totals up to 60,000): the Stirling
    fallback keeps the full tolerance, exact mirror ties included."""
    rng = np.random.default_rng(3)
    t = rng.integers(0, 15000, size=(3000, 4))
    t[::5] = np.stack([t[::5, 0], t[::5, 1], t[::5, 1], t[::5, 0]], axis=1)       # mirror ties
    want = fisher_c.fisher_two_sided(t[:, 0], t[:, 1], t[:, 2], t[:, 3])
    assert _max_rel(twin(t, cap=4096), want) < 1e-11


def test_totals_straddling_the_real_table_cap(twin):
    """Cells of 1-3 million (totals either side of 2^22 = 4,194,304) against the binary128 oracle."""
    rng = np.random.default_rng(12)
    rows = []
    for total in (3_900_000, 4_194_303, 4_194_304, 4_194_305, 4_500_000, 6_000_000):
        for skew in (0.0, 0.001, 0.004):
            n1 = int(total * rng.uniform(0.3, 0.7)); n = int(total * rng.uniform(0.3, 0.7))
            a = int(n1 * n / total * (1.0 + skew))
            rows.append([a, n1 - a, n - a, total - n1 - n + a])
    t = np.array(rows)
    assert (t >= 0).all()
    want = fisher_c.fisher_two_sided(t[:, 0], t[:, 1], t[:, 2], t[:, 3])
    assert (want < 1.0).any() and (want > 1e-300).sum() >= 10
    assert _max_rel(twin(t), want) < 1e-9                      # twin table covers these totals entirely
    assert _max_rel(twin(t, cap=1 << 22), want) < 1e-9         # the kernel's situation: table capped at 2^22


def test_mirror_ties_are_included(twin):
    """Symmetric tables: the mirror term equals pmf(observed) exactly and must be counted."""
    rng = np.random.default_rng(9)
    rows = []
    for _ in range(500):
        a, b = (int(x) for x in rng.integers(0, 3000, 2))
        rows.append([a, b, b, a])
    t = np.array(rows)
    want = fisher_c.fisher_two_sided(t[:, 0], t[:, 1], t[:, 2], t[:, 3])
    assert _max_rel(twin(t), want) < 1e-11


def test_exp_small_against_long_double(twin):
    """The kernel's own exp (table of 2^(j/32) + degree-6 polynomial) over the range of
    log-probabilities: within 1.5 ulp of the exact value (numpy longdouble exp as the reference,
    64-bit significand on x86-64), zero below -746, one final rounding in the subnormal range."""
    rng = np.random.default_rng(7)
    x = np.concatenate([-rng.random(400_000) * 700.0, -rng.random(200_000) * 5.0, rng.random(100_000) * 0.9,
                        -10.0 ** rng.uniform(-12, 0, 100_000), np.array([0.0, -0.0, -1e-300, 0.5, -0.5, -708.0, -745.0])])
    got = twin.exp_small(x)
    want = np.exp(x.astype(np.longdouble))
    rel = np.abs((got.astype(np.longdouble) - want) / want).astype(np.float64)
    assert rel[x > -700.0].max() < 1.5 * 2.0 ** -52
    low = x <= -700.0                                  # towards gradual underflow: one subnormal ulp at most
    assert (np.abs(got[low].astype(np.longdouble) - want[low]) <= np.maximum(want[low] * 2.0 ** -51, 5e-324)).all()
    assert (twin.exp_small(np.array([-746.5, -800.0, -1e308, -np.inf])) == 0.0).all()
    sub = twin.exp_small(np.array([-720.0, -740.0, -745.0]))
    assert (sub > 0).all() and np.allclose(sub, np.exp(np.array([-720.0, -740.0, -745.0])), rtol=1e-6, atol=5e-324)


TIE_RULE_TABLES = [[20000, 19998, 20000, 20002], [20002, 19998, 20000, 20000], [60003, 59997, 60000, 60000]]


def test_tie_rule_is_the_modern_scipy_one(twin):
    """Observed count next to the mode with pmf(observed) within 1e-4 -- but not within 1e-14 -- of
    pmf(mode): scipy >= 1.9 (window 1e-14, stats/_stats_py.py:5082-5086; the image's 1.18.1) sums the
    tails (p ~ 0.994); the scipy 1.4.1 the reference pins returned exactly 1 here (window 1e-4).  The
    kernel arithmetic follows the modern rule, and DESIGN.md section 2 says so."""
    from scipy.stats import fisher_exact
    t = np.array(TIE_RULE_TABLES)
    want = np.array([fisher_exact([[a, b], [c, d]])[1] for a, b, c, d in t.tolist()])
    assert (want < 0.999).all()
    assert _max_rel(twin(t), want) < 1e-9
    np.testing.assert_allclose(twin(t), fisher_c.fisher_two_sided(t[:, 0], t[:, 1], t[:, 2], t[:, 3]), rtol=1e-11)
