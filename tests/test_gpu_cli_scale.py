"""The whole quant command at a size where nothing is hand-checked: 24 synthetic .junc.bed
files x 60,000 junctions through the native readers, the device kernels and the native writers,
against files produced from the oracle with the reference's own per-cell f-strings."""
import argparse
import os

import numpy as np
import pytest

from oracle import oracle_np, ref_port

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_quant_cli_against_oracle_files(tmp_path):
    from splicedice_b200 import quant, synth
    S, J = 24, 60_000
    js = synth.junction_tuples(J, 31)
    counts = synth.counts_host(32, 0, J, S)
    rng = np.random.default_rng(3)
    counts[rng.random(counts.shape) < 0.2] = 0
    man = tmp_path / "manifest.txt"
    with open(man, "w") as m:
        for s in range(S):
            path = tmp_path / f"s{s}.junc.bed"
            order = rng.permutation(J)
            with open(path, "w") as f:
                f.write("".join(f"{js[i][0]}\t{js[i][1]}\t{js[i][2]}\tj\t{counts[i, s]}\t{js[i][3]}\n"
                                for i in order if counts[i, s]))
            m.write(f"s{s}\t{path}\tmeta\tcond{s % 2}\n")
    p = argparse.ArgumentParser()
    quant.add_parser(p)
    quant.run_with(p.parse_args(["-m", str(man), "-o", str(tmp_path / "out"), "--lowCoverageNan", "--minUnique", "6"]))

    # expected files from the oracle: a junction is admitted if ANY sample scores it >= minUnique
    admitted = [i for i in range(J) if (counts[i] >= 6).any()]
    kept = sorted(js[i] for i in admitted)
    row_of = {j: r for r, j in enumerate(kept)}
    mat = np.zeros((len(kept), S), dtype=np.int64)
    for i in admitted:
        mat[row_of[js[i]]] = counts[i]
    arrays = oracle_np.junctions_to_arrays(kept)[:4]
    csr = oracle_np.cluster_csr(*arrays)
    low = (mat > 0) & (mat < 6)                                   # a line exists and scores below minUnique
    ps = oracle_np.ps_f32(mat, csr["row_ptr"], csr["col_idx"], low_mask=low)
    name = lambda j: f"{j[0]}:{j[1]}-{j[2]}:{j[3]}"  # noqa: E731
    names = [name(j) for j in kept]
    header = "cluster\t" + "\t".join(f"s{s}" for s in range(S)) + "\n"
    want_counts = header + "".join(n + "\t" + "\t".join(f"{float(x):.0f}" for x in row) + "\n"
                                   for n, row in zip(names, mat.tolist()))
    want_ps = header + "".join(n + "\t" + "\t".join(f"{x:.3f}" for x in row) + "\n" for n, row in zip(names, ps.tolist()))
    rp, ci = csr["row_ptr"].tolist(), csr["col_idx"].tolist()
    want_clusters = "".join(names[r] + "\t" + ",".join(names[c] for c in ci[rp[r]:rp[r + 1]]) + "\n"
                            for r in range(len(kept)))
    assert open(tmp_path / "out_inclusionCounts.tsv").read() == want_counts
    assert open(tmp_path / "out_allClusters.tsv").read() == want_clusters
    assert open(tmp_path / "out_allPS.tsv").read() == want_ps
    # and the port of the reference's own cluster sweep agrees on a sample of junctions
    adj = ref_port.sweep_clusters(kept)
    for j in kept[::997]:
        assert [name(o) for o in adj[j]] == [names[c] for c in ci[rp[row_of[j]]:rp[row_of[j] + 1]]]
