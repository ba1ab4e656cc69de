"""Drive the UNMODIFIED reference (``/root/reference/splicedice``) in memory.

Test infrastructure, build-container only: ``/root/reference`` does not exist on
the GPU box, so only ``oracle/gen_golden.py`` and the ``needs_reference`` tests
use this module.  Nothing is copied: the modules are imported from where they lie.

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

REFERENCE_ROOT = os.environ.get("SPLICEDICE_REFERENCE", "/root/reference")


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
