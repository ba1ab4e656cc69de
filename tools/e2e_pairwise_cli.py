#!/usr/bin/env python
"""End-to-end `pairwise` CLI on synthetic quant outputs: stage timings of the host mirror
(native table reader, cluster parsing, CSR build, device exclusion sums + Fisher +
Benjamini-Hochberg, device-to-host copy, native str() writer).

    python tools/e2e_pairwise_cli.py [events] [samples] [workdir]
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from splicedice_b200 import junctions as jn, ops, pairwise_fisher, synth, textio  # noqa: E402


def main():
    J = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    work = sys.argv[3] if len(sys.argv) > 3 else tempfile.mkdtemp(prefix="sd_pw_")
    os.makedirs(work, exist_ok=True)
    js = sorted(synth.junction_tuples(J, 11))
    names = [f"{c}:{a}-{b}:{s}" for c, a, b, s in js]
    arrays = synth.junction_arrays(J, 11)
    cl = ops.cluster_build(*arrays[:4])
    rp, ci = cl["row_ptr"].cpu().numpy(), cl["col_idx"].cpu().numpy()
    counts = (synth.counts_host(12, 0, J, S) + synth.counts_host(13, 0, J, S)).astype(np.int32)
    cpath, kpath, opath = (os.path.join(work, n) for n in ("inclusionCounts.tsv", "allClusters.tsv", "pairwise.tsv"))
    textio.write_matrix(cpath, "cluster\t" + "\t".join(f"s{k}" for k in range(S)) + "\n", names, counts)
    with open(kpath, "w") as f:
        f.write("".join(names[r] + "\t" + ",".join(names[c] for c in ci[rp[r]:rp[r + 1]]) + "\n" for r in range(J)))

    t = {}
    t0 = time.perf_counter()
    samples, events, table = pairwise_fisher.getEventCounts(cpath)
    t["read counts table"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    clusters = pairwise_fisher.getClusters(kpath)
    t["read clusters"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    row_ptr, col_idx = jn.csr_from_named_lists(events, clusters, "isin")
    t["CSR from names"] = time.perf_counter() - t0
    torch.cuda.synchronize()
    for rep in range(2):                                 # second pass: warm allocator / table
        t0 = time.perf_counter()
        parray = pairwise_fisher.pairwise_pvalues(events, table, clusters, 0, correction="pairwise")
        torch.cuda.synchronize()
        t["CSR + upload + exclusion sums + Fisher + BH + download"] = time.perf_counter() - t0
    P = parray.shape[1]
    cols = [f"{samples[a]}_{samples[b]}" for a, b in pairwise_fisher.sample_pairs(S)]
    t0 = time.perf_counter()
    textio.write_matrix(opath, "clusterID\t" + "\t".join(cols) + "\n", events, parray, repr_floats=True)
    t["write p-value file"] = time.perf_counter() - t0
    size = os.path.getsize(opath)
    print(f"pairwise CLI stages, {J} events x {S} samples ({P} pairs, {J * P:.3e} p-values, output {size / 1e6:.0f} MB):")
    for k, v in t.items():
        print(f"  {k:58s} {v * 1e3:10.1f} ms")
    # spot check of the file against the matrix
    with open(opath) as f:
        f.readline()
        first = f.readline().rstrip("\n").split("\t")
    assert first[0] == events[0] and [float(x) for x in first[1:]] == parray[0].tolist()
    print("  first row of the file round-trips to the matrix exactly")


if __name__ == "__main__":
    main()
