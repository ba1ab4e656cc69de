"""Synthetic workloads of the shapes BASELINE.json names.

Junction sets: 23 chromosomes x 2 strands, a locus every 60 kb, per locus 4-11
splice sites in a 20 kb window and 2-11 junctions drawn as site pairs at least
50 bp apart (overlap structure: mean degree ~4.5, nnz ~4.5 J).

Counts: an integer-only counter-based generator keyed by (seed, row, col) so that
the device kernel (``sd_synth_counts``) and ``counts_host`` produce the SAME
matrix -- any slab of a 40 GB device matrix can be re-made on the CPU for parity
checks without ever holding the matrix on the host.  The law is an (8-bit
mantissa) approximation of a geometric distribution, i.e. negative_binomial(1, p):
mean ~49 for p = 0.02, matching the read-count scale of SURVEY.md §8d.
"""
from __future__ import annotations

import numpy as np

CHROMS = [f"chr{i}" for i in range(1, 23)] + ["chrX"]
STRANDS = ["+", "-"]

# --- counter-based integer RNG (mirrored bit for bit in csrc/sd_synth.cu) -------
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)

# LOG2_LUT[f] = round(65536 * log2(1 + f/256)), f = 0..255
LOG2_LUT = np.round(65536.0 * np.log2(1.0 + np.arange(256) / 256.0)).astype(np.int64)


def geometric_scale(p: float) -> int:
    """Fixed-point multiplier: value = (L * scale) >> 32 with L = -log2(u) * 65536."""
    return int(round((np.log(2.0) / -np.log1p(-p)) / 65536.0 * 2.0 ** 32))


def _mix64(x):
    x = (x ^ (x >> np.uint64(30))) * _M1
    x = (x ^ (x >> np.uint64(27))) * _M2
    return x ^ (x >> np.uint64(31))


def counts_host(seed: int, row0: int, n_rows: int, n_cols: int, p: float = 0.02,
                rows=None, ld_cols: int | None = None) -> np.ndarray:
    """int32[n_rows, n_cols] block of the synthetic count matrix starting at row0
    (or the listed ``rows``).  ``ld_cols`` = the logical column count of the full
    matrix (the counter is row * ld_cols + col); defaults to n_cols."""
    ld_cols = n_cols if ld_cols is None else ld_cols
    if rows is None:
        rows = np.arange(row0, row0 + n_rows, dtype=np.uint64)
    rows = np.asarray(rows, dtype=np.uint64)
    with np.errstate(over="ignore"):
        ctr = rows[:, None] * np.uint64(ld_cols) + np.arange(n_cols, dtype=np.uint64)[None, :]
        h = _mix64(ctr * _GOLD + np.uint64(seed) * _M2 + np.uint64(1))
    u = (h >> np.uint64(32)).astype(np.int64) | 1           # 32-bit, never zero
    msb = np.floor(np.log2(u.astype(np.float64))).astype(np.int64)   # exact for u < 2^53
    lz = 31 - msb
    norm = (u << lz) & 0xFFFFFFFF                           # leading one at bit 31
    frac = (norm >> 23) & 0xFF
    L = ((lz + 1) << 16) - LOG2_LUT[frac]                   # ~ -log2(u / 2^32) * 65536
    scale = geometric_scale(p)
    return ((L * scale) >> 32).astype(np.int32)


# --- junction sets ---------------------------------------------------------------
def junction_arrays(n_junctions: int, seed: int = 0):
    """Synthetic junction set as arrays (chrom_rank, strand_rank, start, end) with
    ranks under python string order of ``CHROMS`` / ``STRANDS`` (so chr10 < chr2),
    in *random* (set-like) order, plus the name tables.  Exactly n_junctions rows."""
    rng = np.random.default_rng(seed)
    n_cs = len(CHROMS) * len(STRANDS)
    got = np.zeros(0, dtype=np.int64)
    n_loci = int(n_junctions / 5.0 * 1.15) + 64
    base_locus = 0
    while got.size < n_junctions:
        L = n_loci
        locus = np.arange(base_locus, base_locus + L)
        cs = locus % n_cs
        idx = locus // n_cs
        k = rng.integers(4, 12, size=L)
        m = rng.integers(2, 12, size=L)
        sites = rng.integers(0, 20000, size=(L, 11))
        pa = (rng.random((L, 11)) * k[:, None]).astype(np.int64)
        pb = (rng.random((L, 11)) * k[:, None]).astype(np.int64)
        sa = np.take_along_axis(sites, pa, axis=1)
        sb = np.take_along_axis(sites, pb, axis=1)
        lo = np.minimum(sa, sb)
        hi = np.maximum(sa, sb)
        ok = (np.arange(11)[None, :] < m[:, None]) & (hi - lo >= 50)
        # key = cs(6 bits) | locus index(20) | lo site(15) | hi site(15): sorts by (cs, position)
        key = (cs[:, None].astype(np.int64) << 50) | (idx[:, None].astype(np.int64) << 30) | (lo << 15) | hi
        got = np.unique(np.concatenate([got, key[ok]]))
        base_locus += L
        n_loci = max(64, int((n_junctions - got.size) / 4.0) + 64)
    cs = got >> 50
    idx = (got >> 30) & ((1 << 20) - 1)
    if got.size > n_junctions:
        # keep the lowest-numbered loci (loci are dealt round-robin over chrom x strand, so
        # every chromosome/strand stays populated)
        order = np.argsort(idx * n_cs + cs, kind="stable")[:n_junctions]
        got, cs, idx = got[order], cs[order], idx[order]
    perm = rng.permutation(got.size)
    got, cs, idx = got[perm], cs[perm], idx[perm]
    left = (idx * 60000 + 1000 + ((got >> 15) & 0x7FFF)).astype(np.int32)
    right = (idx * 60000 + 1000 + (got & 0x7FFF)).astype(np.int32)
    chrom_id = cs // len(STRANDS)
    strand_id = cs % len(STRANDS)
    chrom_sorted = sorted(CHROMS)
    strand_sorted = sorted(STRANDS)
    chrom_rank = np.array([chrom_sorted.index(c) for c in CHROMS], dtype=np.int32)[chrom_id]
    strand_rank = np.array([strand_sorted.index(s) for s in STRANDS], dtype=np.int32)[strand_id]
    return chrom_rank, strand_rank, left, right, chrom_sorted, strand_sorted


def junction_tuples(n_junctions: int, seed: int = 0):
    """Same set as ``junction_arrays`` as a list of reference-style tuples."""
    c, s, l, r, cn, sn = junction_arrays(n_junctions, seed)
    return [(cn[ci], int(li), int(ri), sn[si]) for ci, si, li, ri in zip(c, s, l, r)]


def adversarial_tuples(seed: int = 0):
    """Touching intervals, deep nests (degree >= 100), duplicate starts, opposite
    strands at identical coordinates, singletons, lexicographic chrom names."""
    rng = np.random.default_rng(seed)
    out = set()
    # touching chain on chr2:+ : [0,100],[100,200],...
    for i in range(50):
        out.add(("chr2", i * 100, (i + 1) * 100, "+"))
    # one huge junction covering 150 small ones on chr10:-
    out.add(("chr10", 1000, 40000, "-"))
    for i in range(150):
        out.add(("chr10", 1100 + i * 200, 1100 + i * 200 + 120, "-"))
    # nested russian dolls on chr1:+
    for i in range(120):
        out.add(("chr1", 5000 + i, 9000 - i, "+"))
    # identical coordinates, both strands
    for i in range(20):
        out.add(("chr1", 20000 + 30 * i, 20500 + 30 * i, "+"))
        out.add(("chr1", 20000 + 30 * i, 20500 + 30 * i, "-"))
    # duplicate starts, different ends
    for i in range(30):
        out.add(("chrX", 700, 800 + 17 * i, "+"))
    # singletons far apart
    for i in range(40):
        out.add(("chr3", 100000 * (i + 1), 100000 * (i + 1) + 77, "-"))
    # random clutter
    for _ in range(300):
        a = int(rng.integers(0, 5000))
        out.add((f"chr{int(rng.integers(1, 23))}", a, a + int(rng.integers(50, 900)), "+-"[int(rng.integers(0, 2))]))
    out = sorted(out)          # set order depends on string hashing: sort before shuffling
    rng.shuffle(out)
    return out
