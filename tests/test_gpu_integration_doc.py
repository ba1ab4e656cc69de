"""INTEGRATION.md is executable: the ctypes stubs it shows a reference maintainer are run here
verbatim (code blocks extracted from the file) on reference-shaped objects and checked against
the oracle."""
import os
import re
import types

import numpy as np
import pytest

from oracle import fisher_c, oracle_np, ref_port

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _blocks():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    return re.findall(r"```python\n(.*?)```", text, flags=re.S)


def test_documented_stubs_run_and_match():
    from splicedice_b200 import native, ops, synth
    ops.require_cuda()
    blocks = _blocks()
    ns = {}
    cwd = os.getcwd()
    os.chdir(ROOT)                                   # the doc loads the library by its in-tree relative path
    try:
        native.load()                                # make sure it is built
        exec(blocks[0], ns)                          # 1. loading
    finally:
        os.chdir(cwd)
    exec(blocks[1], ns)                              # 2. calculatePsi
    exec(blocks[2], ns)                              # 3. getClusters

    js = synth.junction_tuples(3000, 5)
    obj = types.SimpleNamespace()
    obj.junctions = set(js)
    clusters = ns["getClusters"](obj)
    want_clusters = ref_port.sweep_clusters(js)
    assert clusters == want_clusters                 # dict of tuples, reference list order

    obj.clusters = clusters
    obj.junctionIndex = {j: i for i, j in enumerate(sorted(clusters))}
    S = 24
    counts = synth.counts_host(4, 0, len(js), S)
    counts[np.random.default_rng(0).random(counts.shape) < 0.3] = 0
    obj.counts = counts.astype(np.float32)
    obj.low = [(5, 3), (17, 0)]
    obj.args = types.SimpleNamespace(lowCoverageNan=True)
    psi = ns["calculatePsi"](obj)
    want = ref_port.psi_loop(want_clusters, obj.junctionIndex, obj.counts, obj.low)
    np.testing.assert_array_equal(psi.view(np.uint32), want.view(np.uint32))

    # 4. pairwise hot loop
    names = [f"{j[0]}:{j[1]}-{j[2]}:{j[3]}" for j in sorted(clusters)][:300]
    keep = set(names)
    by_name = {f"{j[0]}:{j[1]}-{j[2]}:{j[3]}": [f"{o[0]}:{o[1]}-{o[2]}:{o[3]}" for o in clusters[j]] for j in clusters}
    ns.update(counts=counts[:300, :6].astype(np.float64), events=np.array(names),
              clusters={n: [o for o in by_name[n] if o in keep] for n in names},
              pairs=[(i, j) for i in range(5) for j in range(i + 1, 6)])
    exec(blocks[3], ns)
    pa, pb = oracle_np.all_pairs(6)
    q = fisher_c.pairwise(ns["inc"].astype(np.int64), ns["exc"], pa, pb)
    ok = q > 1e-300
    assert (np.abs(ns["parray"][ok] - q[ok]) / q[ok]).max() < 1e-9

    # multiple-test correction: the documented sd_bh_adjust call, per column
    raw = ns["parray"].copy()
    exec(blocks[4], ns)
    want_adj = np.stack([oracle_np.bh_adjust(raw[:, k]) for k in range(raw.shape[1])], axis=1)
    np.testing.assert_array_equal(ns["corrected"].view(np.uint64), want_adj.view(np.uint64))
