"""K4 parity on the GPU: intron-retention ratio and RSD vs the oracle."""
import numpy as np
import pytest

from oracle import oracle_np
from tests import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ops():
    from splicedice_b200 import ops
    ops.require_cuda()
    return ops


@pytest.mark.parametrize("shape", [(2000, 7), (1500, 128), (900, 1000)])
def test_ir_ratio(shape):
    ops = _ops()
    J, S = shape
    _, csr, counts = util.synthetic_problem(J, S, seed=J)
    rng = np.random.default_rng(J)
    med = rng.poisson(4, size=(J, S)).astype(np.float64)
    med[rng.random((J, S)) < 0.2] = 0.0
    dev = torch.device("cuda", 0)
    c = torch.from_numpy(counts).to(dev)
    m = torch.from_numpy(med).to(dev)
    got = ops.ir_ratio(m, c, csr["row_ptr"], csr["col_idx"]).cpu().numpy()
    want = oracle_np.ir_ratio(med, counts, csr["row_ptr"], csr["col_idx"])
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_array_equal(got[ok], want[ok])
    got1 = ops.ir_ratio(m, c).cpu().numpy()                       # -s: single junction
    want1 = oracle_np.ir_ratio(med, counts, None, None, single_junction=True)
    ok = ~np.isnan(want1)
    np.testing.assert_array_equal(got1[ok], want1[ok])
    np.testing.assert_array_equal(np.isnan(got1), np.isnan(want1))


def test_rsd5():
    ops = _ops()
    rng = np.random.default_rng(2)
    cov = rng.poisson(6, size=(5000, 3, 5)).astype(np.float64)
    cov[::50] = 0.0
    got = ops.rsd5(torch.from_numpy(cov).cuda()).cpu().numpy()
    want = oracle_np.rsd5(cov)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_array_equal(got[ok], want[ok])


@pytest.mark.parametrize("S", [1000, 130, 257])
def test_ir_ratio_fractional_and_extreme_medians(S):
    """The lean ratio epilogue on medians that are not small integers: halves, denormals, huge
    values, inf -- inside the fast divide's range it must equal the IEEE quotient, outside it
    the kernel falls back to the plain divide; ragged last column groups (S % 4 != 0)."""
    ops = _ops()
    J = 1200
    _, csr, counts = util.synthetic_problem(J, S, seed=S)
    rng = np.random.default_rng(S)
    med = rng.poisson(4, size=(J, S)).astype(np.float64) + rng.integers(0, 2, size=(J, S)) * 0.5
    special = np.array([0.0, 5e-324, 1e-310, 3e-200, 2.0 ** -501, 2.0 ** -499, 7.3e10, 2.0 ** 499, 2.0 ** 501, 1e300,
                        1.7976931348623157e308, np.inf, 1 / 3, 1e-17])
    pick = rng.random((J, S)) < 0.15
    med[pick] = rng.choice(special, size=int(pick.sum()))
    counts[rng.random((J, S)) < 0.3] = 0
    buf_m = torch.zeros((J, (S + 3) // 4 * 4), dtype=torch.float64, device="cuda")
    buf_m[:, :S] = torch.from_numpy(med).cuda()
    buf_c = torch.zeros((J, (S + 3) // 4 * 4), dtype=torch.int32, device="cuda")
    buf_c[:, :S] = torch.from_numpy(counts).cuda()
    got = ops.ir_ratio(buf_m[:, :S], buf_c[:, :S], csr["row_ptr"], csr["col_idx"]).cpu().numpy()
    with np.errstate(all="ignore"):
        want = oracle_np.ir_ratio(med, counts, csr["row_ptr"], csr["col_idx"])
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_array_equal(got[ok].view(np.uint64), want[ok].view(np.uint64))
