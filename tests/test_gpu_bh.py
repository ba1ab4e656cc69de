"""Benjamini-Hochberg on the device (sd_bh_adjust) against the numpy restatement of statsmodels'
fdrcorrection (oracle_np.bh_adjust): bit-exact, per column and over the whole matrix."""
import numpy as np
import pytest

from oracle import oracle_np

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ops():
    from splicedice_b200 import native, ops
    ops.require_cuda()
    return native, ops


def _bits(x):
    return np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)


def _want(p, mode):
    if mode == "all":
        return oracle_np.bh_adjust(p.ravel()).reshape(p.shape)
    out = np.empty_like(p)
    for k in range(p.shape[1]):
        out[:, k] = oracle_np.bh_adjust(p[:, k])
    return out


def _pvalues(rng, shape, kind):
    if kind == "uniform":
        return rng.random(shape)
    if kind == "ties":                                  # a few distinct values, many exactly 1.0 (zero-margin tables)
        p = rng.choice(np.array([1.0, 1.0, 0.5, 0.25, 1e-3, 3e-7, 0.0, 0.9999999999999999]), size=shape)
        return p
    p = 10.0 ** (-rng.random(shape) * 300)              # "wide": down to 1e-300 and subnormal results of p * n / k
    p[rng.random(shape) < 0.05] = 5e-324
    return p


@pytest.mark.parametrize("kind", ["uniform", "ties", "wide"])
@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (7, 1), (500, 6), (4096, 3), (4097, 5), (20001, 17), (300, 2016)])
def test_columns_and_all_bit_exact(shape, kind):
    _, ops = _ops()
    rng = np.random.default_rng(shape[0] * 31 + shape[1])
    p = _pvalues(rng, shape, kind)
    d = torch.from_numpy(p).cuda()
    for mode in ("pairwise", "all"):
        got = ops.bh_adjust(d, mode).cpu().numpy()
        assert np.array_equal(_bits(got), _bits(_want(p, mode))), (shape, kind, mode)
    assert np.array_equal(d.cpu().numpy(), p)           # input untouched when out is separate


@pytest.mark.parametrize("kind", ["uniform", "ties", "wide", "fisher"])
@pytest.mark.parametrize("shape", [(8192, 3), (8193, 40), (50_000, 33), (200_000, 7), (524_288, 2), (524_289, 2)])
def test_column_sample_sort_bit_exact(shape, kind, monkeypatch):
    """Column mode from 8,192 rows up runs the per-column sample sort (splitters from a sample,
    equality buckets, windows sorted in shared memory): bit-identical to the oracle and to the
    global-sort path, on uniform values, on a handful of distinct values (every bucket an equality
    bucket, p = 1 filling a third of the column), on values spread over 300 decades, and on a
    Fisher-like mix (8 % exact ones, the rest beta-distributed with heavy ties at small values);
    524,289 rows is the first shape past the sample sort's range."""
    _, ops = _ops()
    rng = np.random.default_rng(shape[0] + shape[1])
    if kind == "fisher":
        p = np.minimum(1.0, rng.beta(0.4, 1.0, size=shape))
        p[rng.random(shape) < 0.08] = 1.0
        small = rng.random(shape) < 0.1
        p[small] = np.round(p[small], 2)                      # ties away from the splitters
    else:
        p = _pvalues(rng, shape, kind)
    d = torch.from_numpy(p).cuda()
    want = _want(p, "pairwise")
    got = ops.bh_adjust(d, "pairwise").cpu().numpy()
    assert np.array_equal(_bits(got), _bits(want)), (shape, kind)
    monkeypatch.setenv("SD_BH_GLOBAL_SORT", "1")
    other = ops.bh_adjust(d, "pairwise").cpu().numpy()
    assert np.array_equal(_bits(other), _bits(want))
    monkeypatch.delenv("SD_BH_GLOBAL_SORT")
    monkeypatch.setenv("SD_BH_TEST_FAIL", "1")                # the oversize-bucket flag: repeat with the global sort
    again = ops.bh_adjust(d, "pairwise").cpu().numpy()
    assert np.array_equal(_bits(again), _bits(want))
    monkeypatch.delenv("SD_BH_TEST_FAIL")
    work = d.clone()                                          # in place
    ops.bh_adjust(work, "pairwise", out=work)
    assert np.array_equal(_bits(work.cpu().numpy()), _bits(want))


def test_column_sample_sort_nan_and_padded_views():
    _, ops = _ops()
    rng = np.random.default_rng(77)
    p = rng.random((20_000, 5))
    p[rng.integers(0, 20_000, 30), 1] = np.nan               # one column with NaNs: the whole column turns NaN-ish like numpy
    p[:, 3] = 1.0
    p[:, 4] = 0.0
    with np.errstate(invalid="ignore"):
        want = _want(p, "pairwise")
    buf = torch.full((20_000, 8), -3.0, dtype=torch.float64, device="cuda")
    buf[:, 2:7] = torch.from_numpy(p).cuda()
    view = buf[:, 2:7]
    got = ops.bh_adjust(view, "pairwise").cpu().numpy()
    for k in range(5):
        assert np.array_equal(np.isnan(got[:, k]), np.isnan(want[:, k])), k
        ok = ~np.isnan(want[:, k])
        assert np.array_equal(_bits(got[ok, k]), _bits(want[ok, k])), k
    assert (buf[:, :2] == -3.0).all() and (buf[:, 7:] == -3.0).all()


def test_long_segments_cross_many_chunks():
    """One segment of 3,000,001 values: > 256 chunks, so the per-segment carry scan loops."""
    _, ops = _ops()
    rng = np.random.default_rng(9)
    p = rng.random((3_000_001, 1)) ** 3
    d = torch.from_numpy(p).cuda()
    want = oracle_np.bh_adjust(p.ravel())
    for mode in ("pairwise", "all"):
        assert np.array_equal(_bits(ops.bh_adjust(d, mode).cpu().numpy().ravel()), _bits(want))
    q = rng.random((150_000, 21))                       # "all" over a 2-D matrix spanning many chunks
    assert np.array_equal(_bits(ops.bh_adjust(torch.from_numpy(q).cuda(), "all").cpu().numpy()), _bits(_want(q, "all")))


def test_in_place_padded_rows_and_views():
    _, ops = _ops()
    rng = np.random.default_rng(4)
    p = rng.random((3000, 10))
    buf = torch.full((3000, 16), -7.0, dtype=torch.float64, device="cuda")
    buf[:, 3:13] = torch.from_numpy(p).cuda()
    view = buf[:, 3:13]                                 # ld 16, base offset by 3 elements
    ops.bh_adjust(view, "pairwise", out=view)
    got = buf.cpu().numpy()
    assert np.array_equal(_bits(got[:, 3:13]), _bits(_want(p, "pairwise")))
    assert (got[:, :3] == -7.0).all() and (got[:, 13:] == -7.0).all()


def test_nan_propagates_like_numpy_minimum():
    _, ops = _ops()
    p = np.array([[0.2, 0.01], [np.nan, 0.5], [0.03, 0.04]])
    with np.errstate(invalid="ignore"):
        want = _want(p, "pairwise")
    got = ops.bh_adjust(torch.from_numpy(p).cuda(), "pairwise").cpu().numpy()
    assert np.isnan(got[:, 0]).all() and np.isnan(want[:, 0]).all()
    assert np.array_equal(_bits(got[:, 1]), _bits(want[:, 1]))


def test_on_fisher_output_and_host_mirror():
    """The matrix the Fisher kernel writes, corrected in place, and the 1-D host-array helper."""
    from splicedice_b200 import pairwise_fisher, synth
    _, ops = _ops()
    J, S = 4000, 10
    c, s, st, en, _, _ = synth.junction_arrays(J, 5)
    cl = ops.cluster_build(c, s, st, en, device=0)
    inc = ops.synth_counts(8, 0, J, S, device=0)[:, :S]
    exc = ops.quant_ps(inc, cl["row_ptr"], cl["col_idx"], want_f32=False, want_exc=True)["exc"]
    pa, pb = ops.all_pairs(S)
    p = ops.fisher_pairwise(inc, exc, pa, pb)
    raw = p.cpu().numpy()
    ops.bh_adjust(p, "pairwise", out=p)
    assert np.array_equal(_bits(p.cpu().numpy()), _bits(_want(raw, "pairwise")))
    assert np.array_equal(_bits(pairwise_fisher.fdr_bh(raw[:, 3])), _bits(oracle_np.bh_adjust(raw[:, 3])))
    np.testing.assert_allclose(pairwise_fisher.fdr_bh([0.01, 0.04, 0.03, 0.005]), [0.02, 0.04, 0.04, 0.02])


def test_argument_errors():
    native, ops = _ops()
    p = torch.rand((10, 4), dtype=torch.float64, device="cuda")
    ws = torch.empty(64, dtype=torch.uint8, device="cuda")
    with pytest.raises(native.NativeCallError) as e:
        native.call("sd_bh_adjust", 10, 4, native.ptr(p), 4, native.ptr(p), 4, 0, native.ptr(ws), 64, None)
    assert e.value.code == 3                            # SD_ERR_WORKSPACE
    with pytest.raises(native.NativeCallError):
        native.call("sd_bh_adjust", 10, 4, native.ptr(p), 3, native.ptr(p), 4, 0, native.ptr(ws), 64, None)
    with pytest.raises(native.NativeCallError):
        native.call("sd_bh_adjust", 10, 4, native.ptr(p), 4, native.ptr(p), 4, 2, native.ptr(ws), 64, None)
    with pytest.raises(TypeError):
        ops.bh_adjust(p.float(), "all")
    with pytest.raises(ValueError):
        ops.bh_adjust(p, "rows")
