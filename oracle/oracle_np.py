"""numpy restatement of the SpliceDICE quant / pairwise arithmetic (TEST ORACLE).

Test infrastructure -- see ``oracle/__init__.py``.  Every function cites the
reference lines it restates (paths relative to ``/root/reference/splicedice``).
All integer work is int64, all floating point follows the reference's dtype
chain exactly so that results are comparable bit for bit.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "rank_strings", "junctions_to_arrays", "cluster_csr", "adjacency_dict",
    "exclusion_sums", "ps_f32", "ps_f64", "ir_ratio", "rsd5", "bh_adjust",
    "all_pairs",
]


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def rank_strings(values):
    """Dense rank of python strings under python's ``str`` ordering.

    The reference sorts tuples whose first / last fields are python strings
    (SPLICEDICE.py:96, :237), so ``chr10 < chr2`` and ``'+' < '-'``.
    """
    uniq = sorted(set(values))
    lut = {v: i for i, v in enumerate(uniq)}
    return np.fromiter((lut[v] for v in values), dtype=np.int32, count=len(values)), uniq


def junctions_to_arrays(junctions):
    """list[(chrom, left, right, strand)] -> int32 arrays (chrom_rank, strand_rank, start, end)."""
    junctions = list(junctions)
    chrom, chrom_names = rank_strings([j[0] for j in junctions])
    strand, strand_names = rank_strings([j[3] for j in junctions])
    start = np.array([j[1] for j in junctions], dtype=np.int32)
    end = np.array([j[2] for j in junctions], dtype=np.int32)
    return chrom, strand, start, end, chrom_names, strand_names


def _inverse_perm(p):
    inv = np.empty_like(p)
    inv[p] = np.arange(p.size, dtype=p.dtype)
    return inv


# --------------------------------------------------------------------------
# a1/a2: overlap adjacency ("clusters") + output row order
# --------------------------------------------------------------------------
def cluster_csr(chrom, strand, start, end):
    """Closed form of ``SPLICEDICE.getClusters`` (SPLICEDICE.py:230-255) plus the
    row index of SPLICEDICE.py:96.

    Within a (chrom, strand) segment sorted by (start, end) -- the sort key of
    SPLICEDICE.py:237 -- junction i and a later junction k are adjacent iff
    ``end_i >= start_k`` (SPLICEDICE.py:250; closed interval).  Because starts
    are sorted, the later neighbours of i are the contiguous run i+1 .. ub_i.
    The reference's list for junction k holds its priors most-recent-first
    (SPLICEDICE.py:247-254 prepend the current junction to the survivor list)
    followed by its laters in ascending order (SPLICEDICE.py:251 appends).

    Returns a dict of int arrays; "pos" = position in cluster order, "row" =
    output row (rank under the tuple order (chrom, left, right, strand)):

    cluster_order[pos] -> input index          out_row[input index] -> row
    row_of_pos[pos]    -> row                  pos_of_row[row]      -> pos
    comp_id[pos]       -> overlap component (contiguous in cluster order)
    row_ptr[J+1], col_idx[nnz]  CSR over rows, columns are rows, list order as
                                the reference's ``clusters[junction]``.
    """
    chrom = np.asarray(chrom, dtype=np.int64)
    strand = np.asarray(strand, dtype=np.int64)
    start = np.asarray(start, dtype=np.int64)
    end = np.asarray(end, dtype=np.int64)
    J = chrom.size
    idx_t = np.int64
    if J == 0:
        z = np.zeros(0, dtype=np.int32)
        return dict(cluster_order=z, out_row=z, row_of_pos=z, pos_of_row=z, comp_id=z,
                    row_ptr=np.zeros(1, dtype=np.int32), col_idx=z, n_comp=0)

    cluster_order = np.lexsort((end, start, strand, chrom)).astype(idx_t)   # :237
    out_order = np.lexsort((strand, end, start, chrom)).astype(idx_t)       # :96 (tuple sort)
    out_row = _inverse_perm(out_order)
    row_of_pos = out_row[cluster_order]
    pos_of_row = _inverse_perm(row_of_pos)

    c = chrom[cluster_order]; s = strand[cluster_order]
    st = start[cluster_order]; en = end[cluster_order]
    seg_head = np.ones(J, dtype=bool)
    seg_head[1:] = (c[1:] != c[:-1]) | (s[1:] != s[:-1])                     # :240
    seg_id = np.cumsum(seg_head) - 1

    # ub[i] = last position of i's segment whose start <= end_i
    off = int(min(st.min(), en.min()))
    big = int(max(st.max(), en.max())) - off + 2
    key_start = seg_id * big + (st - off)
    ub = np.searchsorted(key_start, seg_id * big + (en - off), side="right") - 1
    pos = np.arange(J, dtype=idx_t)
    n_later = np.maximum(ub - pos, 0)

    # component heads: segment head, or no earlier junction of the segment reaches start_k
    comp_head = seg_head.copy()
    # segmented exclusive running max of `end` via the (seg_id, end) packing trick
    packed = seg_id * big + (en - off)
    run = np.maximum.accumulate(packed)
    prev = np.empty(J, dtype=np.int64); prev[0] = -1; prev[1:] = run[:-1]
    prev_seg = prev // big
    prev_end = prev % big + off
    reach = (prev_seg == seg_id) & (prev_end >= st) & (~seg_head)
    comp_head |= ~reach
    comp_id = (np.cumsum(comp_head) - 1)

    # edge list (i < k), k in i+1..ub_i
    total = int(n_later.sum())
    src = np.repeat(pos, n_later)
    first = np.cumsum(n_later) - n_later
    dst = src + 1 + (np.arange(total, dtype=idx_t) - np.repeat(first, n_later))
    n_prior = np.bincount(dst, minlength=J).astype(idx_t)
    deg = n_prior + n_later
    ptr_pos = np.zeros(J + 1, dtype=idx_t)
    np.cumsum(deg, out=ptr_pos[1:])
    cols_pos = np.empty(2 * total, dtype=idx_t)
    # priors of k: sources sorted descending
    o = np.lexsort((-src, dst))
    pri_dst = dst[o]; pri_src = src[o]
    pfirst = np.cumsum(n_prior) - n_prior
    slot = ptr_pos[pri_dst] + (np.arange(total, dtype=idx_t) - pfirst[pri_dst])
    cols_pos[slot] = pri_src
    # laters of i: ascending
    slot = ptr_pos[src] + n_prior[src] + (np.arange(total, dtype=idx_t) - first[src])
    cols_pos[slot] = dst

    # re-index rows/cols into output-row space
    deg_row = deg[pos_of_row]
    row_ptr = np.zeros(J + 1, dtype=idx_t)
    np.cumsum(deg_row, out=row_ptr[1:])
    col_idx = np.empty(2 * total, dtype=idx_t)
    # copy each pos's list to its row's slot
    row_start_for_pos = row_ptr[row_of_pos]
    within = np.arange(2 * total, dtype=idx_t) - np.repeat(ptr_pos[:-1], deg)
    dest = np.repeat(row_start_for_pos, deg) + within
    col_idx[dest] = row_of_pos[cols_pos]

    i32 = np.int32
    return dict(cluster_order=cluster_order.astype(i32), out_row=out_row.astype(i32),
                row_of_pos=row_of_pos.astype(i32), pos_of_row=pos_of_row.astype(i32),
                comp_id=comp_id.astype(i32), row_ptr=row_ptr.astype(i32),
                col_idx=col_idx.astype(i32), n_comp=int(comp_id[-1]) + 1)


def adjacency_dict(junctions, csr):
    """Rebuild the reference's ``clusters`` dict (tuple -> list[tuple]) from a CSR
    in output-row space, for dict-and-list-order equality tests against
    ``SPLICEDICE.getClusters`` (SPLICEDICE.py:230-255)."""
    junctions = list(junctions)
    out_row = csr["out_row"]
    by_row = [None] * len(junctions)
    for i, j in enumerate(junctions):
        by_row[out_row[i]] = j
    rp, ci = csr["row_ptr"], csr["col_idx"]
    return {by_row[r]: [by_row[c] for c in ci[rp[r]:rp[r + 1]]] for r in range(len(junctions))}


# --------------------------------------------------------------------------
# a4/a5: exclusion sums and PS
# --------------------------------------------------------------------------
def exclusion_sums(counts, row_ptr, col_idx):
    """exc[r, :] = sum over the adjacency list of r of counts[c, :]
    (SPLICEDICE.py:303-305; counts_to_ps.py:64-67; duplicates count twice, as
    the reference's ``+=`` loop does).  int64, exact."""
    from scipy import sparse
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    J = counts.shape[0]
    a = sparse.csr_matrix((np.ones(len(col_idx), dtype=np.int64),
                           np.asarray(col_idx, dtype=np.int64),
                           np.asarray(row_ptr, dtype=np.int64)), shape=(J, J))
    return np.asarray(a @ counts, dtype=np.int64)


def ps_f32(counts, row_ptr, col_idx, low_mask=None, exc=None):
    """``SPLICEDICE.calculatePsi`` (SPLICEDICE.py:297-310): inclusions are float32,
    exclusions accumulate in float64 (:303), the divide is float64 and the store
    rounds to float32 (:299,:306); 0/0 -> NaN; low cells -> NaN (:307-309)."""
    if exc is None:
        exc = exclusion_sums(counts, row_ptr, col_idx)
    inc = np.asarray(counts).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        ps = (inc / (inc + exc.astype(np.float64))).astype(np.float32)
    if low_mask is not None:
        ps[np.asarray(low_mask, dtype=bool)] = np.nan
    return ps


def ps_f64(counts, row_ptr, col_idx, exc=None):
    """``counts_to_ps.writePsValues`` (counts_to_ps.py:58-70): float64 counts,
    ``exclusion = inc + sum``, ``ps = inc / exclusion`` in float64."""
    if exc is None:
        exc = exclusion_sums(counts, row_ptr, col_idx)
    inc = np.asarray(counts).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return inc / (inc + exc.astype(np.float64))


# --------------------------------------------------------------------------
# a8: intron retention ratio + RSD
# --------------------------------------------------------------------------
def ir_ratio(median, counts, row_ptr, col_idx, single_junction=False, exc=None):
    """``ir_table.calculateIR`` (ir_table.py:122-132): ``intronCount = inc + sum
    inc[adj]`` (or inc alone with -s), ``IR = median / (median + intronCount)``
    as python floats (float64); ZeroDivisionError -> NaN (so 0/0 and x/0 -> NaN)."""
    inc = np.asarray(counts).astype(np.float64)
    if not single_junction:
        if exc is None:
            exc = exclusion_sums(counts, row_ptr, col_idx)
        inc = inc + exc.astype(np.float64)
    median = np.asarray(median, dtype=np.float64)
    den = median + inc
    with np.errstate(divide="ignore", invalid="ignore"):
        out = median / den
    out[den == 0] = np.nan
    return out


def rsd5(cov):
    """``np.std(cov) / np.mean(cov)`` over the 5 coverage points (ir_table.py:118-120),
    population std.  cov: float64[..., 5]."""
    cov = np.asarray(cov, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.std(cov, axis=-1) / np.mean(cov, axis=-1)


# --------------------------------------------------------------------------
# pairwise helpers
# --------------------------------------------------------------------------
def all_pairs(n_samples):
    """Pair order of pairwise_fisher.py:142-145."""
    a, b = np.triu_indices(n_samples, k=1)
    return a.astype(np.int32), b.astype(np.int32)


def bh_adjust(p):
    """Benjamini-Hochberg as ``statsmodels.stats.multitest.multipletests(...,
    method='fdr_bh')[1]`` (called at pairwise_fisher.py:185,190).  statsmodels is
    absent from this image; restated from its published ``fdrcorrection``:
    sort ascending, p * n / rank, running minimum from the right, clip at 1."""
    p = np.asarray(p, dtype=np.float64)
    n = p.size
    if n == 0:
        return p.copy()
    order = np.argsort(p, kind="stable")
    ps = p[order]
    ecdf = np.arange(1, n + 1, dtype=np.float64) / float(n)
    adj = ps / ecdf
    adj = np.minimum.accumulate(adj[::-1])[::-1]
    adj[adj > 1] = 1
    out = np.empty_like(adj)
    out[order] = adj
    return out
