"""Third-party pin of the Benjamini-Hochberg oracle.

The reference calls statsmodels ``multipletests(method="fdr_bh")`` (pairwise_fisher.py:185,190);
statsmodels is absent from this image, so ``oracle_np.bh_adjust`` restates its ``fdrcorrection``
(sort ascending, ``p / (rank / n)``, running minimum from the right, clip at 1).  What CAN be
proven here: the restatement agrees with the independent implementation this image does ship,
``scipy.stats.false_discovery_control(method='bh')`` -- which evaluates ``p * (n / rank)``, one
rounding placed differently (both forms round twice) -- to within TWO units in the last place on
every value (measured: 15 % of values differ by one ulp, 0.1 % by two), with the same order,
ties, clipping and NaN-free behaviour; and it is within ONE ulp of the exactly rounded rational
value min_{m>=k} p_(m) n / m (python Fractions).  Bit-identity with statsmodels itself stays
unproven (stated in DESIGN.md section 2); the GPU kernel is bit-identical to this oracle
(tests/test_gpu_bh.py), hence within 2 ulp of scipy too."""
import numpy as np
import pytest
from scipy.stats import false_discovery_control

from oracle import oracle_np


def _ulps(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).view(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float64).view(np.int64)
    return np.abs(a - b)


def _cases():
    rng = np.random.default_rng(20261018)
    yield "uniform", rng.random(20_000)
    yield "tiny", rng.random(5_000) * 10.0 ** rng.integers(-300, 0, 5_000)
    yield "ties", rng.integers(0, 40, 9_000) / 40.0
    yield "ones", np.ones(257)
    yield "zeros_and_ones", rng.integers(0, 2, 999).astype(np.float64)
    yield "single", np.array([0.3])
    yield "fisher_like", np.minimum(1.0, rng.beta(0.3, 1.0, 200_000))
    yield "sorted_desc", np.sort(rng.random(4_096))[::-1].copy()
    yield "two", np.array([0.04, 0.01])


@pytest.mark.parametrize("name,p", list(_cases()), ids=[c[0] for c in _cases()])
def test_bh_oracle_within_one_ulp_of_scipy(name, p):
    got = oracle_np.bh_adjust(p)
    want = false_discovery_control(p, method="bh")
    assert got.shape == want.shape and not np.isnan(got).any()
    assert _ulps(got, want).max() <= 2, f"{name}: {_ulps(got, want).max()} ulp"
    assert (got <= 1.0).all() and (got >= p).all()
    # adjusted values keep the order of the raw ones (ties stay ties)
    order = np.argsort(p, kind="stable")
    assert (np.diff(got[order]) >= 0).all()
    same = np.diff(p[order]) == 0
    assert (np.diff(got[order])[same] == 0).all()


def test_bh_oracle_per_column_equals_scipy_axis0():
    rng = np.random.default_rng(4)
    p = rng.random((3_000, 17))
    got = np.stack([oracle_np.bh_adjust(p[:, k]) for k in range(p.shape[1])], axis=1)
    want = false_discovery_control(p, axis=0, method="bh")
    assert _ulps(got, want).max() <= 2


def test_bh_oracle_within_one_ulp_of_exact_rationals():
    from fractions import Fraction
    rng = np.random.default_rng(9)
    p = np.concatenate([rng.random(700), rng.integers(0, 30, 300) / 30.0])
    n = p.size
    order = np.argsort(p, kind="stable")
    exact = [None] * n
    run = None
    for k in range(n - 1, -1, -1):
        v = Fraction(float(p[order[k]])) * n / (k + 1)
        run = v if run is None or v < run else run
        exact[order[k]] = min(run, Fraction(1))
    want = np.array([float(v) for v in exact])                  # Fraction -> float rounds correctly
    assert _ulps(oracle_np.bh_adjust(p), want).max() <= 1


def test_bh_known_answers():
    """Hand-checked textbook vector: adj_(k) = min_{m >= k} p_(m) n / m."""
    p = np.array([0.01, 0.04, 0.03, 0.005])
    np.testing.assert_allclose(oracle_np.bh_adjust(p), [0.02, 0.04, 0.04, 0.02], rtol=0, atol=1e-18)
    np.testing.assert_array_equal(oracle_np.bh_adjust(np.array([0.5, 0.9, 1.0])), [1.0, 1.0, 1.0])
