"""Shared helpers for the test-suite (inputs + oracle plumbing)."""
import os

import numpy as np

from oracle import oracle_np
from splicedice_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_junction_arrays(g):
    """(chrom_rank, strand_rank, start, end) of a golden quant fixture, ranks under python str order."""
    chrom, _ = oracle_np.rank_strings([str(c) for c in g["chrom"]])
    strand, _ = oracle_np.rank_strings([str(s) for s in g["strand"]])
    return chrom, strand, g["start"].astype(np.int32), g["end"].astype(np.int32)


def golden_row_order(g):
    """Golden quant fixtures store counts in ``sorted(junctions)`` order; returns the output row of
    every stored junction (identity when the fixture is already sorted)."""
    js = list(zip([str(c) for c in g["chrom"]], g["start"].tolist(), g["end"].tolist(), [str(s) for s in g["strand"]]))
    order = sorted(range(len(js)), key=lambda i: js[i])
    out_row = np.empty(len(js), dtype=np.int64)
    out_row[order] = np.arange(len(js))
    return out_row


def synthetic_problem(n_junctions, n_samples, seed, zero_frac=0.2):
    c, s, st, en, _, _ = synth.junction_arrays(n_junctions, seed)
    csr = oracle_np.cluster_csr(c, s, st, en)
    counts = synth.counts_host(seed + 1, 0, n_junctions, n_samples)
    rng = np.random.default_rng(seed + 2)
    counts[rng.random(counts.shape) < zero_frac] = 0
    return (c, s, st, en), csr, counts


def bits32(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def bits64(x):
    return np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)
