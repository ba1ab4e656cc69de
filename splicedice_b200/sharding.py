"""Row-slab sharding of the quant / pairwise work across GPUs (host logic, numpy only).

Adjacency never leaves an overlap component (SPLICEDICE.py:240-254: the sweep resets per
(chrom, strand) and drops priors that stop overlapping), so the count matrix can be cut into
contiguous output-row slabs wherever no adjacency edge crosses the cut.  Each GPU then owns its
slab of ``counts`` plus the matching CSR slice with columns rebased to the slab -- no halo, no
exchange step, results return by host gather (or an optional NCCL all-gather of PS rows).

A cut before row r is *safe* iff every row below r has all of its neighbours below r.  With
``reach[i] = max(i, max(col_idx of row i))`` that is ``max(reach[:r]) < r``, a prefix maximum.
"""
from __future__ import annotations

import numpy as np


def safe_cuts(row_ptr, col_idx) -> np.ndarray:
    """bool[J + 1]: safe[r] is True when rows [0, r) and [r, J) share no adjacency edge."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    col_idx = np.asarray(col_idx, dtype=np.int64)
    J = row_ptr.size - 1
    safe = np.ones(J + 1, dtype=bool)
    if J == 0:
        return safe
    reach = np.arange(J, dtype=np.int64)
    deg = np.diff(row_ptr)
    if col_idx.size:
        has = deg > 0
        # np.maximum.reduceat needs non-empty segments: reduce over rows that have neighbours
        seg_max = np.maximum.reduceat(col_idx, row_ptr[:-1][has])
        reach[has] = np.maximum(reach[has], seg_max)
    prefix = np.maximum.accumulate(reach)
    safe[1:J] = prefix[:J - 1] < np.arange(1, J)
    return safe


def row_weights(row_ptr, n_samples: int = 1) -> np.ndarray:
    """Work per row: one own-row read + one read per adjacency entry, times the row length."""
    deg = np.diff(np.asarray(row_ptr, dtype=np.int64))
    return (1 + deg) * max(int(n_samples), 1)


def partition_rows(row_ptr, col_idx, n_shards: int, weights=None):
    """Cut [0, J) into ``n_shards`` contiguous slabs at safe cuts, balancing ``weights``
    (default: 1 + degree, i.e. nnz-balanced).  Returns a list of (row_begin, row_end); slabs may
    be empty when there are fewer safe cuts than shards."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    J = row_ptr.size - 1
    if n_shards < 1:
        raise ValueError("n_shards must be >= 1")
    if weights is None:
        weights = row_weights(row_ptr)
    weights = np.asarray(weights, dtype=np.float64)
    if weights.size != J:
        raise ValueError("weights must have one entry per row")
    safe = np.flatnonzero(safe_cuts(row_ptr, col_idx))          # sorted cut positions, has 0 and J
    cum = np.concatenate([[0.0], np.cumsum(weights)])
    total = cum[-1]
    cuts = [0]
    for k in range(1, n_shards):
        target = total * k / n_shards
        # nearest safe cut (by cumulative weight) that does not move backwards
        i = np.searchsorted(cum[safe], target)
        cand = [c for c in (i - 1, i) if 0 <= c < safe.size]
        best = min(cand, key=lambda c: abs(cum[safe[c]] - target))
        cuts.append(max(int(safe[best]), cuts[-1]))
    cuts.append(J)
    return [(cuts[k], cuts[k + 1]) for k in range(n_shards)]


def shard_csr(row_ptr, col_idx, row_begin: int, row_end: int):
    """CSR of rows [row_begin, row_end) with columns rebased to the slab.  Raises if an edge
    leaves the slab (the cut was not safe)."""
    row_ptr = np.asarray(row_ptr)
    col_idx = np.asarray(col_idx)
    lo, hi = int(row_ptr[row_begin]), int(row_ptr[row_end])
    cols = col_idx[lo:hi].astype(np.int64) - row_begin
    if cols.size and (cols.min() < 0 or cols.max() >= row_end - row_begin):
        raise ValueError(f"rows [{row_begin}, {row_end}) are not closed under adjacency")
    return (row_ptr[row_begin:row_end + 1].astype(np.int64) - lo).astype(np.int32), cols.astype(np.int32)


def imbalance(parts, weights) -> float:
    """max shard weight / mean shard weight."""
    cum = np.concatenate([[0.0], np.cumsum(np.asarray(weights, dtype=np.float64))])
    w = np.array([cum[b] - cum[a] for a, b in parts])
    return float(w.max() / w.mean()) if w.mean() > 0 else 1.0
