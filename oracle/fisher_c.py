"""ctypes wrapper over oracle/fisher_oracle.c (TEST ORACLE -- see oracle/__init__.py)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import build_oracle

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_oracle.build())
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.fisher_oracle_batch.argtypes = [ctypes.c_int64, i64p, i64p, i64p, i64p,
                                            ctypes.POINTER(ctypes.c_double)]
        lib.fisher_oracle_batch.restype = ctypes.c_int
        lib.fisher_oracle_support.argtypes = [ctypes.c_int64, i64p, i64p, i64p, i64p, i64p]
        lib.fisher_oracle_support.restype = ctypes.c_int
        _lib = lib
    return _lib


def _as_i64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int64).ravel())


def fisher_two_sided(a, b, c, d) -> np.ndarray:
    """Two-sided p of [[a, b], [c, d]] element-wise (scipy.stats.fisher_exact semantics)."""
    shape = np.broadcast(a, b, c, d).shape
    a, b, c, d = (_as_i64(np.broadcast_to(x, shape)) for x in (a, b, c, d))
    out = np.empty(a.size, dtype=np.float64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    rc = _load().fisher_oracle_batch(a.size, a.ctypes.data_as(i64p), b.ctypes.data_as(i64p),
                                     c.ctypes.data_as(i64p), d.ctypes.data_as(i64p),
                                     out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    if rc != 0:
        raise ValueError("fisher oracle: negative entry or allocation failure")
    return out.reshape(shape)


def support_size(a, b, c, d) -> np.ndarray:
    shape = np.broadcast(a, b, c, d).shape
    a, b, c, d = (_as_i64(np.broadcast_to(x, shape)) for x in (a, b, c, d))
    out = np.empty(a.size, dtype=np.int64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    _load().fisher_oracle_support(a.size, a.ctypes.data_as(i64p), b.ctypes.data_as(i64p),
                                  c.ctypes.data_as(i64p), d.ctypes.data_as(i64p),
                                  out.ctypes.data_as(i64p))
    return out.reshape(shape)


def pairwise(inc, exc, pair_a, pair_b) -> np.ndarray:
    """p[j, k] for tables [[inc[j,a_k], inc[j,b_k]], [exc[j,a_k], exc[j,b_k]]]
    (pairwise_fisher.py:164-165)."""
    inc = np.asarray(inc, dtype=np.int64)
    exc = np.asarray(exc, dtype=np.int64)
    return fisher_two_sided(inc[:, pair_a], inc[:, pair_b], exc[:, pair_a], exc[:, pair_b])
