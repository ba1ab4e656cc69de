"""The CPU oracle against fixtures made by EXECUTING the reference (oracle/gen_golden.py).

These run everywhere (no GPU, no reference tree): they pin the oracle that the GPU parity tests
trust.  The reference ships no tests of its own, so reference-run outputs
are the pin.
"""
import json
import os

import numpy as np
import pytest

from oracle import fisher_c, oracle_np, ref_port
from tests import util


@pytest.mark.parametrize("name", ["quant_adversarial.npz", "quant_synth_3k.npz"])
def test_cluster_csr_equals_reference_adjacency(name):
    g = util.load_npz(name)
    csr = oracle_np.cluster_csr(*util.golden_junction_arrays(g))
    np.testing.assert_array_equal(csr["out_row"], util.golden_row_order(g))
    np.testing.assert_array_equal(csr["row_ptr"], g["row_ptr"])
    np.testing.assert_array_equal(csr["col_idx"], g["col_idx"])          # list order included
    # components: every adjacency edge stays inside one component; ids are contiguous in cluster order
    comp_of_row = np.empty(len(csr["comp_id"]), dtype=np.int64)
    comp_of_row[csr["row_of_pos"]] = csr["comp_id"]
    rows = np.repeat(np.arange(len(comp_of_row)), np.diff(csr["row_ptr"]))
    assert np.all(comp_of_row[rows] == comp_of_row[csr["col_idx"]])
    assert np.all(np.diff(csr["comp_id"]) >= 0)


@pytest.mark.parametrize("name", ["quant_adversarial.npz", "quant_synth_3k.npz"])
def test_ps_f32_bits_equal_reference(name):
    g = util.load_npz(name)
    counts = g["counts"]
    ps = oracle_np.ps_f32(counts, g["row_ptr"], g["col_idx"])
    np.testing.assert_array_equal(util.bits32(ps), g["psi_nolow_bits"])
    low = np.zeros(counts.shape, dtype=bool)
    low[g["low"][:, 0], g["low"][:, 1]] = True
    ps = oracle_np.ps_f32(counts, g["row_ptr"], g["col_idx"], low_mask=low)
    np.testing.assert_array_equal(util.bits32(ps), g["psi_bits"])


def test_ref_port_matches_golden():
    g = util.load_npz("quant_adversarial.npz")
    js = list(zip([str(c) for c in g["chrom"]], g["start"].tolist(), g["end"].tolist(), [str(s) for s in g["strand"]]))
    adj = ref_port.sweep_clusters(js)
    index = ref_port.row_index(adj)
    rp, ci = g["row_ptr"], g["col_idx"]
    for j, r in index.items():
        assert [index[o] for o in adj[j]] == ci[rp[r]:rp[r + 1]].tolist()
    low = [tuple(x) for x in g["low"].tolist()]
    ps = ref_port.psi_loop(adj, index, g["counts"].astype(np.float32), low)
    np.testing.assert_array_equal(util.bits32(ps), g["psi_bits"])


def test_fisher_oracle_vs_scipy_golden():
    g = util.load_npz("fisher_tables.npz")
    t, p = g["tables"], g["p"]
    got = fisher_c.fisher_two_sided(t[:, 0], t[:, 1], t[:, 2], t[:, 3])
    ok = p > 1e-300
    rel = np.abs(got[ok] - p[ok]) / p[ok]
    assert rel.max() < 1e-9, rel.max()
    assert np.array_equal(got[p == 1.0], p[p == 1.0])


def test_fisher_oracle_vs_exact_rational():
    g = util.load_npz("fisher_exact_small.npz")
    t, p = g["tables"], g["p"]
    got = fisher_c.fisher_two_sided(t[:, 0], t[:, 1], t[:, 2], t[:, 3])
    np.testing.assert_allclose(got, p, rtol=1e-13, atol=0)


def _read_tsv(path):
    rows = [l.rstrip("\n").split("\t") for l in open(path)]
    return rows[0], rows[1:]


@pytest.mark.parametrize("case", ["survey_vector", "cli_5x_pairwise"])
def test_pairwise_files_from_reference(case, golden_dir):
    """Exclusions via the clusters file + binary128 Fisher + BH == the reference's pairwise output."""
    exp = os.path.join(golden_dir, case, "expected")
    _, crow = _read_tsv(os.path.join(exp, "ref_inclusionCounts.tsv"))
    names = [r[0] for r in crow]
    counts = np.array([[int(x) for x in r[1:]] for r in crow], dtype=np.int64)
    idx = {n: i for i, n in enumerate(names)}
    adj = {}
    for l in open(os.path.join(exp, "ref_allClusters.tsv")):
        f = l.split()
        adj[f[0]] = f[1].split(",") if len(f) > 1 else []
    exc = np.array([counts[[idx[o] for o in adj[n]]].sum(axis=0) if adj[n] else np.zeros(counts.shape[1], int)
                    for n in names])
    pa, pb = oracle_np.all_pairs(counts.shape[1])
    p = fisher_c.pairwise(counts, exc, pa, pb)
    for mode in ("none", "pairwise", "all"):
        _, rows = _read_tsv(os.path.join(exp, f"pairwise_{mode}.tsv"))
        want = np.array([[float(x) for x in r[1:]] for r in rows])
        assert [r[0] for r in rows] == names
        if mode == "none":
            got = p
        elif mode == "pairwise":
            got = np.stack([oracle_np.bh_adjust(p[:, k]) for k in range(p.shape[1])], axis=1)
        else:
            got = oracle_np.bh_adjust(p.ravel()).reshape(p.shape)
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=0)


def test_ir_golden(golden_dir):
    d = os.path.join(golden_dir, "ir_small")
    samples = json.load(open(os.path.join(d, "samples.json")))
    hdr, crow = _read_tsv(os.path.join(d, "counts.tsv"))
    names = [r[0] for r in crow]
    counts = np.array([[int(x) for x in r[1:]] for r in crow], dtype=np.int64)
    idx = {n: i for i, n in enumerate(names)}
    rp = [0]
    ci = []
    adj = {}
    for l in open(os.path.join(d, "clusters.tsv")):
        f = l.rstrip("\n").split("\t")
        adj[f[0]] = [x for x in f[1].split(",") if x] if len(f) > 1 else []
    for n in names:
        ci += [idx[o] for o in adj[n]]
        rp.append(len(ci))
    med = np.zeros(counts.shape)
    cov = np.zeros(counts.shape + (5,))
    for s, smp in enumerate(samples):
        for l in open(os.path.join(d, "cov", f"{smp}_intron_coverage.txt")):
            f = l.rstrip("\n").split("\t")
            i = idx[f"{f[0]}:{f[1]}-{f[2]}:{f[5]}"]
            med[i, s] = float(f[4])
            cov[i, s] = [float(x) for x in f[-1].split(",")]
    ir = oracle_np.ir_ratio(med, counts, np.array(rp), np.array(ci))
    ir1 = oracle_np.ir_ratio(med, counts, None, None, single_junction=True)
    rsd = oracle_np.rsd5(cov)
    for fname, arr in (("ref_intron_retention.tsv", ir), ("ref_single_intron_retention.tsv", ir1),
                       ("ref_intron_retention_RSD.tsv", rsd)):
        h, rows = _read_tsv(os.path.join(d, "expected", fname))
        col = [samples.index(x[:-4] if x.endswith("_RSD") else x) for x in h[1:]]
        for r in rows:
            got = arr[idx[r[0]]][col]
            assert [f"{x:0.03f}" for x in got] == r[1:], (fname, r[0])
