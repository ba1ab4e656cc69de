"""Drive the UNMODIFIED reference in memory.

Test infrastructure / CPU baseline.  The reference's modules are imported from
``/root/reference/splicedice`` where that exists (the build container) and otherwise from
``oracle/_ref/`` -- the git-ignored pip install of the same tree that ``oracle/build_ref.py``
makes and that travels to the GPU box with the snapshot.  ``oracle/gen_golden.py``, the
``needs_reference`` tests and ``bench.py``'s CPU legs use this module; nothing under
``splicedice_b200/`` does.

The reference package's ``__main__`` imports pysam / statsmodels (absent), so the
hot-path modules are imported directly and ``SPLICEDICE`` objects are made with
``object.__new__`` + attribute injection (the constructor runs the whole
file-based pipeline, SPLICEDICE.py:73-131).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for cand in (os.environ.get("SPLICEDICE_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "splicedice", "SPLICEDICE.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "splicedice", "SPLICEDICE.py"))


def _bh_stub():
    """statsmodels stand-in for pairwise_fisher.py:120 (import inside run_with)."""
    from . import oracle_np

    def multipletests(pvals, alpha=0.05, method="fdr_bh"):
        assert method == "fdr_bh"
        adj = oracle_np.bh_adjust(np.asarray(pvals, dtype=float))
        return adj <= alpha, adj, None, None

    sm = types.ModuleType("statsmodels")
    st = types.ModuleType("statsmodels.stats")
    mt = types.ModuleType("statsmodels.stats.multitest")
    mt.multipletests = multipletests
    sm.stats = st
    st.multitest = mt
    return {"statsmodels": sm, "statsmodels.stats": st, "statsmodels.stats.multitest": mt}


def load(name: str):
    """Import ``splicedice.<name>`` from the reference tree (harness-only shims)."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    try:
        import statsmodels  # noqa: F401
    except Exception:
        for k, v in _bh_stub().items():
            sys.modules.setdefault(k, v)
    if not hasattr(np, "float"):
        np.float = float  # ir_table.py:118 uses the alias removed in numpy 1.24
    return importlib.import_module(f"splicedice.{name}")


class _Args:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def quant_args(**over):
    d = dict(maxLength=50000, minLength=50, minOverhang=5, drim=False, noMultimap=False,
             filter="gtag_only", minUnique=5, lowCoverageNan=False, minEntropy=1.0)
    d.update(over)
    return _Args(**d)


def ref_get_clusters(junctions):
    """``SPLICEDICE.getClusters`` on an in-memory junction set."""
    mod = load("SPLICEDICE")
    obj = object.__new__(mod.SPLICEDICE)
    obj.junctions = set(junctions)
    return obj.getClusters()


def ref_calculate_psi(junctions, counts_f32, low=None):
    """``getClusters`` + row index (SPLICEDICE.py:96) + ``calculatePsi`` on in-memory
    data.  counts rows must already be in ``sorted(junctions)`` order."""
    mod = load("SPLICEDICE")
    obj = object.__new__(mod.SPLICEDICE)
    obj.junctions = set(junctions)
    obj.clusters = obj.getClusters()
    obj.junctionIndex = {j: i for i, j in enumerate(sorted(obj.clusters))}
    obj.manifest = [None] * counts_f32.shape[1]
    obj.counts = np.asarray(counts_f32, dtype=np.float32)
    obj.low = list(low) if low is not None else []
    obj.args = quant_args(lowCoverageNan=low is not None)
    return obj.clusters, obj.calculatePsi()


def reset_sample_state():
    """``Sample.sampleList`` / ``groups`` are class attributes (SPLICEDICE.py:13-14)."""
    mod = load("SPLICEDICE")
    mod.Sample.sampleList = []
    mod.Sample.groups = {}


# ------------------------------------------------------------------------------------------
# timed CPU baseline: the reference's own methods, optionally fanned out over processes
# ------------------------------------------------------------------------------------------
def closed_row_slabs(rows_sorted, n_slabs):
    """Cut ``sorted(junctions)`` (the reference's output-row order, SPLICEDICE.py:96) into about
    ``n_slabs`` contiguous row ranges that no overlap can cross: a cut before row k is allowed when
    the chromosome changes or every earlier junction of that chromosome (either strand) ends before
    row k starts.  Each range is then a self-contained junction set for getClusters/calculatePsi
    whose rows are a contiguous block of the count matrix."""
    n = len(rows_sorted)
    if n == 0:
        return []
    chrom_id = np.zeros(n, dtype=np.int64)
    left = np.fromiter((j[1] for j in rows_sorted), dtype=np.int64, count=n)
    right = np.fromiter((j[2] for j in rows_sorted), dtype=np.int64, count=n)
    names = [j[0] for j in rows_sorted]
    change = np.array([i > 0 and names[i] != names[i - 1] for i in range(n)])
    chrom_id = np.cumsum(change)
    # running max of `right` within the chromosome: offset every chromosome into its own band
    band = (int(right.max()) + 2)
    run = np.maximum.accumulate(right + chrom_id * band)
    ok = np.ones(n, dtype=bool)
    ok[1:] = (run[:-1] < left[1:] + chrom_id[1:] * band)
    cuts = np.flatnonzero(ok)
    want = np.linspace(0, n, n_slabs + 1)[1:-1]
    picked = sorted({int(cuts[min(np.searchsorted(cuts, w), len(cuts) - 1)]) for w in want} - {0})
    edges = [0] + picked + [n]
    return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]


_FORK = {}


def _slab_task(k):
    """getClusters -> row index -> calculatePsi of one closed row slab, all three the reference's
    own code (SPLICEDICE.py:230-255, :96, :297-310) on an object made with object.__new__."""
    mod = _FORK["mod"]
    np.seterr(all="ignore")                     # numpy's 0/0 warning (SPLICEDICE.py:306), once per row otherwise
    r0, r1 = _FORK["slabs"][k]
    obj = object.__new__(mod.SPLICEDICE)
    obj.junctions = set(_FORK["rows"][r0:r1])
    obj.clusters = obj.getClusters()
    obj.junctionIndex = {junction: i for i, junction in enumerate(sorted(obj.clusters))}
    obj.manifest = _FORK["manifest"]
    obj.counts = _FORK["counts"][r0:r1]
    obj.low = []
    obj.args = _FORK["args"]
    psi = obj.calculatePsi()
    keep = _FORK.get("keep")
    if keep is not None:
        keep[r0:r1] = psi
    with np.errstate(invalid="ignore"):
        return r0, r1, float(np.nansum(psi, dtype=np.float64)), int(np.isnan(psi).sum())


class RefQuantPool:
    """The reference's quant compute pass on in-memory data, timed per step.

    ``workers == 1``: one object, the whole problem, one core (the reference is single-threaded).
    ``workers > 1``: the problem is cut into closed row slabs (above) that forked worker processes
    pull from a queue; every slab still runs the reference's own three stages.  The workers are
    forked ONCE, when the pool is made (inputs are inherited through fork, nothing is pickled), so a
    timed step is the slab work alone, not process start-up.  ``counts_f32`` rows are in
    ``sorted(junctions)`` order."""

    def __init__(self, rows_sorted, counts_f32, workers=1, keep=None):
        self.workers = max(1, int(workers))
        S = counts_f32.shape[1]
        slabs = [(0, len(rows_sorted))] if self.workers == 1 else closed_row_slabs(rows_sorted, self.workers * 6)
        _FORK.clear()
        _FORK.update(mod=load("SPLICEDICE"), rows=rows_sorted, counts=counts_f32, slabs=slabs, manifest=[None] * S,
                     args=quant_args(), keep=keep if self.workers == 1 else None)
        self.n_slabs = len(slabs)
        self.pool = None
        if self.workers > 1:
            import multiprocessing as mp
            with warnings_off():
                self.pool = mp.get_context("fork").Pool(self.workers)

    def step(self):
        """(seconds, checksum, nan_count) of one getClusters + row index + calculatePsi pass."""
        import time
        t0 = time.perf_counter()
        if self.pool is None:
            parts = [_slab_task(0)]
        else:
            parts = self.pool.map(_slab_task, range(self.n_slabs), chunksize=1)
        dt = time.perf_counter() - t0
        return dt, float(sum(p[2] for p in parts)), int(sum(p[3] for p in parts))

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None
        _FORK.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def ref_quant_seconds(rows_sorted, counts_f32, workers=1, keep=None):
    """One timed step of a fresh RefQuantPool: (seconds, checksum, nan_count); ``keep``
    (float32[J, S], workers == 1 only) receives the PS matrix."""
    with warnings_off(), RefQuantPool(rows_sorted, counts_f32, workers, keep) as pool:
        return pool.step()


class warnings_off:
    """numpy's 0/0 RuntimeWarning (SPLICEDICE.py:306) once per row would flood stderr."""

    def __enter__(self):
        import warnings
        self._ctx = warnings.catch_warnings()
        self._ctx.__enter__()
        warnings.simplefilter("ignore")
        self._err = np.seterr(all="ignore")
        return self

    def __exit__(self, *exc):
        np.seterr(**self._err)
        return self._ctx.__exit__(*exc)


def ref_pairwise_tests_per_second(n_events, n_samples, seed, workdir):
    """tests/s of the unmodified ``pairwise_fisher.run_with`` (correction "none") on ``n_events``
    events x all sample pairs: the files it reads are written first (outside the timed region), the
    timed region is run_with itself -- file parsing, the np.isin / sum / fisher_exact loop
    (pairwise_fisher.py:154-180) and the writer."""
    import contextlib
    import io
    import time
    rng = np.random.default_rng(seed)
    inc = rng.negative_binomial(2, 0.02, size=(n_events, n_samples))
    names = [f"chr1:{100 * i}-{100 * i + 150}:+" for i in range(n_events)]
    counts_tsv = os.path.join(workdir, "counts.tsv")
    clusters_tsv = os.path.join(workdir, "clusters.tsv")
    with open(counts_tsv, "w") as f:
        f.write("cluster\t" + "\t".join(f"s{k}" for k in range(n_samples)) + "\n")
        for i, name in enumerate(names):
            f.write(name + "\t" + "\t".join(str(int(x)) for x in inc[i]) + "\n")
    with open(clusters_tsv, "w") as f:
        for i, name in enumerate(names):
            f.write(name + "\t" + ",".join(names[j] for j in (i - 1, i + 1) if 0 <= j < n_events) + "\n")
    mod = load("pairwise_fisher")
    args = _Args(inclusionSPLICEDICE=counts_tsv, clusters=clusters_tsv, chi2=False, multiple_test_correction="none",
                 filter_list=None, output=os.path.join(workdir, "pairwise.tsv"))
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        mod.run_with(args)
    dt = time.perf_counter() - t0
    return n_events * (n_samples * (n_samples - 1) // 2) / dt
