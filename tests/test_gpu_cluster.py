"""K1 parity on the GPU: adjacency CSR (list order included), components and row order, bit-exact."""
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


def _compare(arrays):
    ops = _ops()
    want = oracle_np.cluster_csr(*arrays)
    got = ops.cluster_build(*arrays)
    torch.cuda.synchronize()
    assert got["nnz"] == int(want["row_ptr"][-1])
    assert got["n_comp"] == want["n_comp"]
    for k in ("cluster_order", "out_row", "row_of_pos", "comp_id", "row_ptr", "col_idx"):
        np.testing.assert_array_equal(got[k].cpu().numpy(), want[k], err_msg=k)
    return got, want


@pytest.mark.parametrize("name", ["quant_adversarial.npz", "quant_synth_3k.npz"])
def test_golden_adjacency_from_reference(name):
    """CSR built on the device == the adjacency lists SPLICEDICE.getClusters produced (list order too)."""
    g = util.load_npz(name)
    arrays = util.golden_junction_arrays(g)
    got, _ = _compare(arrays)
    # golden rows are stored in sorted(junctions) order = output row order of the stored junctions
    out_row = util.golden_row_order(g)
    rp, ci = g["row_ptr"], g["col_idx"]
    grp, gci = got["row_ptr"].cpu().numpy(), got["col_idx"].cpu().numpy()
    dev_row = got["out_row"].cpu().numpy()
    for i in range(len(out_row)):
        r = dev_row[i]
        assert r == out_row[i]
    np.testing.assert_array_equal(grp, rp)
    np.testing.assert_array_equal(gci, ci)


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (50, 2), (5000, 3), (120000, 4)])
def test_synthetic_sets(n, seed):
    from splicedice_b200 import synth
    c, s, st, en, _, _ = synth.junction_arrays(n, seed)
    _compare((c, s, st, en))


def test_adversarial_structures():
    from splicedice_b200 import synth
    js = synth.adversarial_tuples(7)
    c, s, st, en, _, _ = oracle_np.junctions_to_arrays(js)
    _compare((c, s, st, en))


def test_empty_set():
    ops = _ops()
    z = np.zeros(0, dtype=np.int32)
    got = ops.cluster_build(z, z, z, z)
    assert got["nnz"] == 0 and got["n_comp"] == 0 and got["row_ptr"].cpu().numpy().tolist() == [0]


@pytest.mark.parametrize("seed", range(12))
def test_random_interval_sets(seed):
    """Touching, nested and duplicate-start intervals on several chromosomes / strands, sizes 1..3000:
    device CSR == oracle CSR (list order, components, row order)."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 3000))
    chroms = ["chr1", "chr10", "chr2", "chrX", "GL000219.1"]
    out = set()
    while len(out) < n:
        a = int(rng.integers(0, 5 * n))
        kind = int(rng.integers(0, 4))
        length = [int(rng.integers(1, 40)), int(rng.integers(40, 3000)), 50, int(rng.integers(1, 5))][kind]
        out.add((chroms[int(rng.integers(0, len(chroms)))], a, a + length, "+-"[int(rng.integers(0, 2))]))
    js = sorted(out)
    rng.shuffle(js)
    arrays = oracle_np.junctions_to_arrays(js)[:4]
    _compare(arrays)


def test_one_giant_component():
    """A junction spanning everything: every other junction is its neighbour (degree J - 1)."""
    js = [("chr1", 0, 10_000_000, "+")] + [("chr1", 100 * i + 10, 100 * i + 60, "+") for i in range(20000)]
    arrays = oracle_np.junctions_to_arrays(js)[:4]
    got, want = _compare(arrays)
    assert got["n_comp"] == 1 and got["nnz"] == 2 * 20000


@pytest.mark.parametrize("seed", range(6))
def test_dense_small_coordinates_and_three_strand_labels(seed):
    """Junction sets drawn from a handful of coordinates (touching, nested, identical-start /
    identical-end intervals and opposite-strand twins everywhere), chromosome names whose string
    order differs from their numeric order and a third strand label ('.', which BED allows) --
    the structure tests/test_reference_live.py pins against the reference with hypothesis."""
    rng = np.random.default_rng(100 + seed)
    chroms, strands = ["chr1", "chr11", "chr2", "MT", "1"], ["+", "-", "."]
    n = int(rng.integers(1, 400))
    out = set()
    while len(out) < n:
        a = int(rng.integers(0, 30))
        out.add((chroms[int(rng.integers(0, 5))], a, a + int(rng.integers(1, 12)), strands[int(rng.integers(0, 3))]))
    js = list(out)
    rng.shuffle(js)
    got, want = _compare(oracle_np.junctions_to_arrays(js)[:4])
    # the oracle's adjacency is the brute-force closed-interval overlap on (chromosome, strand)
    adj = oracle_np.adjacency_dict(js, want)
    for j, lst in adj.items():
        brute = {k for k in js if k != j and k[0] == j[0] and k[3] == j[3] and k[1] <= j[2] and j[1] <= k[2]}
        assert set(lst) == brute and len(lst) == len(brute)


@pytest.mark.parametrize("run", [2, 64, 65, 400])
def test_runs_of_junctions_sharing_a_start(run):
    """The cluster order comes from ONE sort by (chrom, strand, start) plus a fix-up of the runs that
    share all three; runs of up to 64 are sorted in place, a longer one makes the build repeat with
    the two-pass sort.  Both sides of that limit, several runs per set, against the oracle."""
    rng = np.random.default_rng(run)
    js = set()
    for k, (chrom, strand) in enumerate([("chr1", "+"), ("chr1", "-"), ("chr2", "+")]):
        start = 1000 + 7 * k
        for e in rng.choice(np.arange(start + 50, start + 50 + 3 * run), size=run, replace=False):
            js.add((chrom, start, int(e), strand))
        for _ in range(40):                                        # clutter around the runs
            a = int(rng.integers(900, 1400))
            js.add((chrom, a, a + int(rng.integers(50, 400)), strand))
    js = list(js)
    rng.shuffle(js)
    c, s, st, en, _, _ = oracle_np.junctions_to_arrays(js)
    _compare((c, s, st, en))
