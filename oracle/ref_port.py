"""Loop-for-loop CPU port of the reference's hot loops (TEST ORACLE / CPU BASELINE).

Test infrastructure -- see ``oracle/__init__.py``.  ``bench.py`` times the
UNMODIFIED reference (installed into ``oracle/_ref`` by ``oracle/build_ref.py``);
this port is what it falls back to (``kind: "port"``) when no reference tree is
installed.  It keeps the reference's data structures (dict of tuples -> list of
tuples, one numpy vector op per adjacency entry, one ``scipy.stats.fisher_exact``
call per table) so that it costs what the reference costs;
``tests/test_reference_live.py`` checks it against the reference itself (bitwise
PS, adjacency dict with list order, pairwise loop) wherever a reference tree
exists.
"""
from __future__ import annotations

import numpy as np


def sweep_clusters(junctions):
    """Port of ``SPLICEDICE.getClusters`` (SPLICEDICE.py:230-255).

    One pass over the junctions sorted by (chrom, strand, start, end) keeping the
    list of earlier junctions that can still overlap; overlap is closed
    (``prior_end >= start``), the list is rebuilt newest-first each step and a
    prior that stops overlapping is never looked at again.
    """
    adjacency = {}
    live = []
    where = (None, None)
    for cur in sorted(junctions, key=lambda j: (j[0], j[3], j[1], j[2])):
        if (cur[0], cur[3]) != where:
            where = (cur[0], cur[3])
            live = []
        mine = adjacency[cur] = []
        keep = [cur]
        cur_start = cur[1]
        for old in live:
            if old[2] >= cur_start:
                adjacency[old].append(cur)
                mine.append(old)
                keep.append(old)
        live = keep
    return adjacency


def row_index(adjacency):
    """Port of SPLICEDICE.py:96 -- rows follow python's tuple order."""
    return {j: r for r, j in enumerate(sorted(adjacency))}


def psi_loop(adjacency, index, counts_f32, low=None):
    """Port of ``SPLICEDICE.calculatePsi`` (SPLICEDICE.py:297-310): float32
    inclusions, a fresh float64 exclusion vector per junction, one vector add per
    adjacency entry, float64 divide stored to float32, then the low-cell NaNs."""
    n_rows, n_samples = counts_f32.shape
    out = np.zeros((n_rows, n_samples), dtype="float32")
    with np.errstate(divide="ignore", invalid="ignore"):
        for junction in sorted(adjacency):
            r = index[junction]
            inc = counts_f32[r, :]
            exc = np.zeros(n_samples)
            for other in adjacency[junction]:
                exc += counts_f32[index[other], :]
            out[r, :] = inc / (inc + exc)
    if low is not None:
        for r, s in low:
            out[r, s] = np.nan
    return out


def psi_rows(adjacency, index, counts_f32, keys):
    """The body of ``psi_loop`` for the listed junctions only, returning just their rows
    (row-slab fan-out of the CPU baseline over processes)."""
    n_samples = counts_f32.shape[1]
    out = np.zeros((len(keys), n_samples), dtype="float32")
    with np.errstate(divide="ignore", invalid="ignore"):
        for k, junction in enumerate(keys):
            inc = counts_f32[index[junction], :]
            exc = np.zeros(n_samples)
            for other in adjacency[junction]:
                exc += counts_f32[index[other], :]
            out[k, :] = inc / (inc + exc)
    return out


def ps_from_tables(clusters_by_name, counts_by_name):
    """Port of the arithmetic in ``counts_to_ps.writePsValues`` (counts_to_ps.py:61-68),
    returning {name: float64[S]} instead of writing the TSV."""
    out = {}
    with np.errstate(divide="ignore", invalid="ignore"):
        for name, overlaps in clusters_by_name.items():
            total = counts_by_name[name].copy()
            for other in overlaps:
                if other == "":
                    continue
                total += counts_by_name[other]
            out[name] = counts_by_name[name] / total
    return out


def pairwise_loop(events, counts_f64, clusters_by_name, pairs=None, progress=None):
    """Port of the hot loop of ``pairwise_fisher.run_with`` (pairwise_fisher.py:142-180):
    per event the exclusion vector is the sum of the rows whose *name* is in the
    event's cluster (``np.isin`` on strings, :158-160), then one two-sided
    ``scipy.stats.fisher_exact`` per sample pair (:165,:179)."""
    from scipy.stats import fisher_exact

    events = np.asarray(events)
    n_samples = counts_f64.shape[1]
    if pairs is None:
        pairs = [(i, j) for i in range(n_samples - 1) for j in range(i + 1, n_samples)]
    table_p = []
    for n, inc in enumerate(counts_f64):
        if progress is not None and n % 50 == 0:
            progress(n)
        members = counts_f64[np.isin(events, clusters_by_name[events[n]])]
        exc = np.sum(members, axis=0)
        row = []
        for a, b in pairs:
            row.append(fisher_exact([[inc[a], inc[b]], [exc[a], exc[b]]])[1])
        table_p.append(row)
    return np.array(table_p, dtype=np.float64).reshape(len(events), len(pairs))


def ir_loop(median, counts, adjacency_rows, single_junction=False):
    """Port of the arithmetic of ``ir_table.calculateIR`` (ir_table.py:122-132) on
    dense arrays: python-float accumulation per (junction, sample)."""
    n_rows, n_samples = counts.shape
    out = np.empty((n_rows, n_samples), dtype=np.float64)
    for r in range(n_rows):
        for s in range(n_samples):
            total = float(counts[r, s])
            if not single_junction:
                for c in adjacency_rows[r]:
                    total += float(counts[c, s])
            m = float(median[r, s])
            try:
                out[r, s] = m / (m + total)
            except ZeroDivisionError:
                out[r, s] = np.nan
    return out
