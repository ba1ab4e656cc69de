"""Junction bookkeeping shared by the host-side commands: names, string ranks, CSR <-> dict.

A junction is the reference's tuple ``(chrom, left, right, strand)`` and its text name is
``chrom:left-right:strand`` (SPLICEDICE.py:312-314).  The device works on integer arrays; the
only string work left on the host is ranking chromosome / strand names under python's ``str``
order (the order every ``sorted()`` of the reference uses: SPLICEDICE.py:96,237).
"""
from __future__ import annotations

import numpy as np


def junction_name(j) -> str:
    return f"{j[0]}:{j[1]}-{j[2]}:{j[3]}"


def parse_name(name: str):
    """``chrom:left-right:strand`` -> tuple (counts_to_ps.py:53-56).  ValueError if malformed."""
    chrom, coords, strand = name.split(":")
    left, right = (int(x) for x in coords.split("-"))
    return (chrom, left, right, strand)


def dense_ranks(values):
    """(int32 ranks, sorted unique values) of python strings under ``str`` ordering."""
    uniq = sorted(set(values))
    lut = {v: i for i, v in enumerate(uniq)}
    return np.fromiter((lut[v] for v in values), dtype=np.int32, count=len(values)), uniq


class JunctionTable:
    """Column-wise copy of a junction collection, in the caller's order."""

    def __init__(self, junctions):
        self.tuples = list(junctions)
        n = len(self.tuples)
        self.chrom_rank, self.chrom_names = dense_ranks([j[0] for j in self.tuples])
        self.strand_rank, self.strand_names = dense_ranks([j[3] for j in self.tuples])
        self.start = np.fromiter((j[1] for j in self.tuples), dtype=np.int64, count=n)
        self.end = np.fromiter((j[2] for j in self.tuples), dtype=np.int64, count=n)
        if n and (self.start.min() < 0 or self.end.max() >= 2 ** 31 or self.start.max() >= 2 ** 31 or self.end.min() < 0):
            raise ValueError("junction coordinates must lie in [0, 2^31)")
        if len(self.chrom_names) >= 1 << 25 or len(self.strand_names) >= 1 << 8:
            raise ValueError("too many distinct chromosome / strand names for the device sort key")
        self.start = self.start.astype(np.int32)
        self.end = self.end.astype(np.int32)

    def __len__(self):
        return len(self.tuples)

    def arrays(self):
        return self.chrom_rank, self.strand_rank, self.start, self.end


def rows_in_output_order(tuples, out_row):
    """list of junction tuples indexed by output row."""
    by_row = [None] * len(tuples)
    for i, r in enumerate(np.asarray(out_row).tolist()):
        by_row[r] = tuples[i]
    return by_row


def adjacency_dict(by_row, row_ptr, col_idx):
    """The reference's ``clusters`` dict (tuple -> list of tuples, reference list order) from a
    CSR in output-row space."""
    rp = np.asarray(row_ptr).tolist()
    ci = np.asarray(col_idx).tolist()
    return {by_row[r]: [by_row[c] for c in ci[rp[r]:rp[r + 1]]] for r in range(len(by_row))}


def csr_from_named_lists(names, clusters, mode):
    """CSR over ``names`` (row order = ``names``) from a dict name -> list of names.

    mode "sum" : every listed name contributes once per listing; empty strings are skipped; a
                 listed name that has no row raises KeyError -- the loop of
                 counts_to_ps.writePsValues (counts_to_ps.py:63-67).
    mode "isin": set semantics of ``counts[np.isin(events, clusters[event])]``
                 (pairwise_fisher.py:158): duplicates count once, names without a row are
                 ignored, and a name that labels several rows selects all of them.  A row whose
                 own name has no entry in ``clusters`` raises KeyError, as the reference does.
    """
    first = {}
    every = {}
    for i, n in enumerate(names):
        first.setdefault(n, i)
        every.setdefault(n, []).append(i)
    row_ptr = np.zeros(len(names) + 1, dtype=np.int64)
    cols = []
    for i, n in enumerate(names):
        listed = clusters[n]                      # KeyError propagates, as in the reference
        if mode == "sum":
            for o in listed:
                if o == "":
                    continue
                if o not in first:
                    raise KeyError(o)
                cols.append(first[o])
        elif mode == "isin":
            mine = []
            for o in set(listed):
                mine.extend(every.get(o, ()))
            cols.extend(sorted(mine))
        else:
            raise ValueError(mode)
        row_ptr[i + 1] = len(cols)
    if row_ptr[-1] >= 2 ** 31:
        raise OverflowError("adjacency does not fit int32 indices")
    return row_ptr.astype(np.int32), np.asarray(cols, dtype=np.int32)
