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
