"""The BASELINE.json shapes at full size, checked through properties that do not need the whole
matrix on the host: rows sampled from the device result are re-made on the CPU with the
counter-based generator (splicedice_b200.synth) and compared bit for bit with the oracle;
ranges / NaN rules / symmetry are checked on the whole device result."""
import numpy as np
import pytest

from oracle import fisher_c, oracle_np
from tests import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _rows_check(ops, synth, seed, J, S, row0, csr, ps, n_rows=160):
    rp, ci = csr
    rng = np.random.default_rng(seed)
    rows = np.sort(rng.choice(J, size=n_rows, replace=False))
    need = sorted(set(rows.tolist()) | {int(c) for r in rows for c in ci[rp[r]:rp[r + 1]]})
    host = synth.counts_host(seed, 0, len(need), S, rows=[row0 + r for r in need], ld_cols=S).astype(np.int64)
    at = {r: k for k, r in enumerate(need)}
    got = ps[torch.from_numpy(rows).to(ps.device)].cpu().numpy()
    for k, r in enumerate(rows):
        inc = host[at[int(r)]]
        exc = host[[at[int(c)] for c in ci[rp[r]:rp[r + 1]]]].sum(axis=0) if rp[r + 1] > rp[r] else np.zeros(S, np.int64)
        with np.errstate(divide="ignore", invalid="ignore"):
            want = (inc.astype(np.float32) / (inc.astype(np.float32) + exc.astype(np.float64))).astype(np.float32)
        np.testing.assert_array_equal(util.bits32(got[k]), util.bits32(want), err_msg=f"row {r}")


@pytest.mark.parametrize("J,S,label", [(400_000, 1000, "configs[1] GTEx-scale"),
                                       (125_000, 10_000, "configs[3] TCGA-scale, one of 8 shards")])
def test_quant_ps_full_size(J, S, label):
    from splicedice_b200 import ops, synth
    ops.require_cuda()
    dev = torch.device("cuda", 0)
    cl = ops.cluster_build(*synth.junction_arrays(J, 77)[:4])
    csr = (cl["row_ptr"].cpu().numpy(), cl["col_idx"].cpu().numpy())
    counts = ops.synth_counts(5, 0, J, S, device=dev)
    ps = ops.quant_ps(counts, cl["row_ptr"], cl["col_idx"])["ps_f32"]
    torch.cuda.synchronize()
    _rows_check(ops, synth, 5, J, S, 0, csr, ps)
    # whole-matrix properties: 0 <= PS <= 1 or NaN; NaN exactly where the junction and all of its
    # neighbours are zero; a junction without neighbours has PS 1 wherever it is covered
    finite = ~torch.isnan(ps)
    assert bool(((ps[finite] >= 0) & (ps[finite] <= 1)).all())
    lonely = torch.from_numpy(np.flatnonzero(np.diff(csr[0]) == 0)).to(dev)
    sub_c, sub_p = counts[lonely], ps[lonely]
    assert bool(torch.isnan(sub_p[sub_c == 0]).all()) and bool((sub_p[sub_c > 0] == 1).all())
    assert bool((ps[counts == 0][~torch.isnan(ps[counts == 0])] == 0).all())
    # idempotence / determinism: a second launch writes the same bits
    again = ops.quant_ps(counts, cl["row_ptr"], cl["col_idx"])["ps_f32"]
    assert torch.equal(ps.view(torch.int32), again.view(torch.int32))


def test_pairwise_fisher_full_size():
    """configs[2]: 64 samples (2,016 pairs) x 200,000 junctions = 4.03e8 tests."""
    from splicedice_b200 import ops, synth
    ops.require_cuda()
    dev = torch.device("cuda", 0)
    J, S = 200_000, 64
    cl = ops.cluster_build(*synth.junction_arrays(J, 78)[:4])
    inc = ops.synth_counts(8, 0, J, S, device=dev) + ops.synth_counts(9, 0, J, S, device=dev)
    exc = ops.quant_ps(inc, cl["row_ptr"], cl["col_idx"], want_f32=False, want_exc=True)["exc"]
    pa, pb = ops.all_pairs(S)
    p = ops.fisher_pairwise(inc, exc, pa, pb)
    torch.cuda.synchronize()
    assert p.shape == (J, 2016)
    assert not bool(torch.isnan(p).any()) and bool(((p >= 0) & (p <= 1)).all())    # p < 1e-308 underflows to 0, as scipy
    # zero-margin rows (no neighbours, so exc = 0) are exactly 1
    lonely = torch.from_numpy(np.flatnonzero(np.diff(cl["row_ptr"].cpu().numpy()) == 0)).to(dev)
    assert bool((p[lonely] == 1).all())
    # symmetry: swapping the two samples of every pair leaves p unchanged
    rows = torch.from_numpy(np.sort(np.random.default_rng(1).choice(J, 4000, replace=False))).to(dev)
    swapped = ops.fisher_pairwise(inc[rows].contiguous(), exc[rows].contiguous(), pb, pa)
    torch.testing.assert_close(swapped, p[rows], rtol=1e-12, atol=0)
    # sampled rows against the binary128 oracle
    sel = rows[:150]
    want = fisher_c.pairwise(inc[sel].cpu().numpy(), exc[sel].cpu().numpy(), pa, pb)
    got = p[sel].cpu().numpy()
    ok = want > 1e-300
    assert (np.abs(got[ok] - want[ok]) / want[ok]).max() < 1e-11


def test_ir_ratio_full_size():
    """configs[4]: intron-retention ratio from precomputed coverage, 1,000 samples x 400,000 junctions."""
    from splicedice_b200 import ops, synth
    ops.require_cuda()
    dev = torch.device("cuda", 0)
    J, S = 400_000, 1000
    cl = ops.cluster_build(*synth.junction_arrays(J, 79)[:4])
    rp, ci = cl["row_ptr"].cpu().numpy(), cl["col_idx"].cpu().numpy()
    counts = ops.synth_counts(6, 0, J, S, device=dev)
    med = (ops.synth_counts(7, 0, J, S, p=0.2, device=dev) % 9).double()          # coverage medians 0..8
    ir = ops.ir_ratio(med, counts, cl["row_ptr"], cl["col_idx"])
    torch.cuda.synchronize()
    rows = np.sort(np.random.default_rng(2).choice(J, 120, replace=False))
    need = sorted(set(rows.tolist()) | {int(c) for r in rows for c in ci[rp[r]:rp[r + 1]]})
    at = {r: k for k, r in enumerate(need)}
    host = synth.counts_host(6, 0, len(need), S, rows=need, ld_cols=S).astype(np.int64)
    med_h = (synth.counts_host(7, 0, len(rows), S, p=0.2, rows=rows, ld_cols=S) % 9).astype(np.float64)
    got = ir[torch.from_numpy(rows).to(dev)].cpu().numpy()
    for k, r in enumerate(rows):
        total = host[at[int(r)]] + (host[[at[int(c)] for c in ci[rp[r]:rp[r + 1]]]].sum(axis=0) if rp[r + 1] > rp[r] else 0)
        den = med_h[k] + total
        with np.errstate(divide="ignore", invalid="ignore"):
            want = np.where(den == 0, np.nan, med_h[k] / den)
        np.testing.assert_array_equal(np.isnan(got[k]), np.isnan(want))
        np.testing.assert_array_equal(got[k][~np.isnan(want)], want[~np.isnan(want)])
    finite = ~torch.isnan(ir)
    assert bool(((ir[finite] >= 0) & (ir[finite] <= 1)).all())


def test_bh_full_size_properties_and_sampled_columns():
    """The per-pair Benjamini-Hochberg correction of a configs[2]-sized p-value matrix
    (200,000 x 2,016: the per-column sample sort): 24 sampled columns bit-identical to the oracle,
    and on the whole matrix the properties that hold for every column -- adjusted >= raw, <= 1,
    order-preserving (a column sorted by p has non-decreasing adjusted values), exact ones stay
    ones, idempotent bits on a second call, and the same bits from the global-sort path."""
    import os
    from splicedice_b200 import ops
    ops.require_cuda()
    J, P = 200_000, 2016
    g = torch.Generator(device="cuda").manual_seed(3)
    p = torch.rand((J, P), dtype=torch.float64, device="cuda", generator=g) ** 3
    p[torch.rand((J, P), device="cuda", generator=g) < 0.08] = 1.0
    adj = ops.bh_adjust(p, "pairwise")
    assert bool((adj >= p).all()) and bool((adj <= 1.0).all())
    cols = np.sort(np.random.default_rng(1).choice(P, size=24, replace=False))
    host_p = p[:, torch.from_numpy(cols).cuda()].cpu().numpy()
    host_a = adj[:, torch.from_numpy(cols).cuda()].cpu().numpy()
    for k in range(len(cols)):
        want = oracle_np.bh_adjust(host_p[:, k])
        np.testing.assert_array_equal(util.bits64(host_a[:, k]), util.bits64(want), err_msg=f"column {cols[k]}")
    order = torch.argsort(p[:, :64], dim=0, stable=True)
    assert bool((torch.diff(torch.gather(adj[:, :64], 0, order), dim=0) >= 0).all())
    again = ops.bh_adjust(p, "pairwise")
    assert torch.equal(adj.view(torch.int64), again.view(torch.int64))
    del again
    os.environ["SD_BH_GLOBAL_SORT"] = "1"
    try:
        other = ops.bh_adjust(p, "pairwise")
    finally:
        del os.environ["SD_BH_GLOBAL_SORT"]
    assert torch.equal(adj.view(torch.int64), other.view(torch.int64))


def test_cluster_build_full_size():
    """configs[3]'s junction set (1,000,000 junctions) on the device against the numpy oracle, bit for
    bit: cluster order, output rows, components, CSR."""
    from splicedice_b200 import ops, synth
    ops.require_cuda()
    arrays = synth.junction_arrays(1_000_000, 20261021)[:4]
    want = oracle_np.cluster_csr(*arrays)
    got = ops.cluster_build(*arrays)
    assert got["nnz"] == int(want["row_ptr"][-1]) and got["n_comp"] == want["n_comp"]
    for k in ("cluster_order", "out_row", "row_of_pos", "comp_id", "row_ptr", "col_idx"):
        np.testing.assert_array_equal(got[k].cpu().numpy(), want[k], err_msg=k)
