"""The oracle against the UNMODIFIED reference executed live (build container only: needs
/root/reference; skipped elsewhere).  Property-style: random junction sets with touching,
nested and duplicate-start intervals, both strands, lexicographic chromosome names."""
import numpy as np
import pytest

from oracle import fisher_c, oracle_np, ref_harness, ref_port

pytestmark = pytest.mark.needs_reference


def _random_junctions(rng, n):
    chroms = ["chr1", "chr10", "chr2", "chrX"]
    out = set()
    while len(out) < n:
        a = int(rng.integers(0, 400))
        kind = rng.integers(0, 4)
        length = [int(rng.integers(1, 40)), int(rng.integers(40, 300)), 50, int(rng.integers(1, 5))][kind]
        out.add((chroms[int(rng.integers(0, 4))], a, a + length, "+-"[int(rng.integers(0, 2))]))
    js = list(out)
    rng.shuffle(js)
    return js


@pytest.mark.parametrize("seed", range(25))
def test_clusters_and_psi_match_the_reference(seed):
    rng = np.random.default_rng(seed)
    js = _random_junctions(rng, int(rng.integers(1, 160)))
    S = int(rng.integers(1, 9))
    counts = rng.integers(0, 60, size=(len(js), S))
    counts[rng.random(counts.shape) < 0.3] = 0
    order = sorted(range(len(js)), key=lambda i: js[i])
    counts_sorted = counts[order].astype(np.float32)                 # rows in sorted(junctions) order
    low = [(int(r), int(s)) for r, s in zip(rng.integers(0, len(js), 5), rng.integers(0, S, 5))]
    clusters, psi = ref_harness.ref_calculate_psi(js, counts_sorted, low)
    # numpy oracle
    arrays = oracle_np.junctions_to_arrays(js)[:4]
    csr = oracle_np.cluster_csr(*arrays)
    assert oracle_np.adjacency_dict(js, csr) == clusters                # list order included
    mask = np.zeros(counts.shape, bool)
    for r, s in low:
        mask[r, s] = True
    ps = oracle_np.ps_f32(counts_sorted.astype(np.int64), csr["row_ptr"], csr["col_idx"], low_mask=mask)
    np.testing.assert_array_equal(ps.view(np.uint32), psi.view(np.uint32))
    # loop-for-loop port (the timed CPU baseline)
    adj = ref_port.sweep_clusters(js)
    assert adj == clusters
    index = ref_port.row_index(adj)
    np.testing.assert_array_equal(ref_port.psi_loop(adj, index, counts_sorted, low).view(np.uint32), psi.view(np.uint32))


def test_pairwise_loop_port_matches_the_reference(tmp_path):
    """ref_port.pairwise_loop and the binary128 oracle against pairwise_fisher.run_with."""
    rng = np.random.default_rng(3)
    js = sorted(_random_junctions(rng, 40))
    counts = rng.integers(0, 80, size=(len(js), 4))
    clusters = ref_harness.ref_get_clusters(js)
    name = lambda j: f"{j[0]}:{j[1]}-{j[2]}:{j[3]}"  # noqa: E731
    ctsv, ltsv, out = tmp_path / "c.tsv", tmp_path / "l.tsv", tmp_path / "p.tsv"
    ctsv.write_text("cluster\ta\tb\tc\td\n" + "".join(name(j) + "\t" + "\t".join(str(x) for x in counts[i]) + "\n"
                                                     for i, j in enumerate(js)))
    ltsv.write_text("".join(name(j) + "\t" + ",".join(name(o) for o in clusters[j]) + "\n" for j in js))
    mod = ref_harness.load("pairwise_fisher")
    args = ref_harness._Args(inclusionSPLICEDICE=str(ctsv), clusters=str(ltsv), chi2=False,
                             multiple_test_correction="none", filter_list=None, output=str(out))
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        mod.run_with(args)
    want = np.array([[float(x) for x in l.split("\t")[1:]] for l in out.read_text().splitlines()[1:]])
    names = [name(j) for j in js]
    by_name = {name(j): [name(o) for o in clusters[j]] for j in js}
    got = ref_port.pairwise_loop(names, counts.astype(float), by_name)
    np.testing.assert_array_equal(got, want)
    idx = {n: i for i, n in enumerate(names)}
    exc = np.array([counts[[idx[o] for o in by_name[n]]].sum(axis=0) if by_name[n] else np.zeros(4, int) for n in names])
    pa, pb = oracle_np.all_pairs(4)
    q = fisher_c.pairwise(counts, exc, pa, pb)
    np.testing.assert_allclose(q, want, rtol=1e-9, atol=0)


def test_fisher_oracle_against_scipy_live():
    from scipy.stats import fisher_exact
    rng = np.random.default_rng(11)
    t = np.concatenate([rng.integers(0, 12, size=(400, 4)), rng.integers(0, 400, size=(300, 4))])
    want = np.array([fisher_exact([[a, b], [c, d]])[1] for a, b, c, d in t.tolist()])
    got = fisher_c.fisher_two_sided(t[:, 0], t[:, 1], t[:, 2], t[:, 3])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=0)
    assert np.array_equal(got == 1.0, want == 1.0)


def test_adjacency_property_on_dense_small_coordinates():
    """hypothesis: junction sets drawn from a handful of coordinates (so touching, nested,
    identical-start, identical-end and opposite-strand twins abound), chromosome names whose
    string order differs from their numeric order, and a third strand label ('.', which BED
    allows): the oracle's adjacency dict -- list order included -- and its component labels against
    the reference's sweep; exclusion sums against a brute-force closed-interval overlap count."""
    from hypothesis import given, settings, strategies as st

    junction = st.tuples(st.sampled_from(["chr1", "chr11", "chr2", "MT", "1"]), st.integers(0, 12), st.integers(1, 9),
                         st.sampled_from(["+", "-", "."])).map(lambda t: (t[0], t[1], t[1] + t[2], t[3]))

    @settings(max_examples=120, deadline=None)
    @given(st.sets(junction, min_size=1, max_size=40), st.randoms(use_true_random=False))
    def check(jset, rnd):
        js = list(jset)
        rnd.shuffle(js)
        clusters = ref_harness.ref_get_clusters(js)
        csr = oracle_np.cluster_csr(*oracle_np.junctions_to_arrays(js)[:4])
        assert oracle_np.adjacency_dict(js, csr) == clusters
        assert ref_port.sweep_clusters(js) == clusters
        # closed-interval overlap on the same chromosome and strand, brute force
        for j, adj in clusters.items():
            want = {k for k in js if k != j and k[0] == j[0] and k[3] == j[3] and k[1] <= j[2] and j[1] <= k[2]}
            assert set(adj) == want and len(adj) == len(want)

    check()
