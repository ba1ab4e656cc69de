"""Row-text writers over ``sd_host_format_rows`` (native, multi-threaded, byte-identical to the
reference's per-cell f-strings)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import native

_KINDS = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2}


def format_rows(matrix, names=None, threads=0, repr_floats=False) -> bytes:
    """b"name\\tv\\tv...\\n" for every row of a 2-D float32 / float64 ('.3f') or int32 matrix;
    ``repr_floats``: float64 as python's str(p) (the pairwise writer)."""
    m = np.asarray(matrix)
    if m.ndim != 2 or m.dtype not in _KINDS or (repr_floats and m.dtype != np.float64):
        raise TypeError("format_rows: 2-D float32 / float64 / int32 matrix expected")
    if m.strides[1] != m.itemsize:
        m = np.ascontiguousarray(m)
    rows, cols = m.shape
    if rows == 0:
        return b""
    blob = off = None
    if names is not None:
        enc = [n.encode() for n in names]
        if len(enc) != rows:
            raise ValueError("format_rows: one name per row expected")
        off = np.zeros(rows + 1, dtype=np.int64)
        np.cumsum([len(e) for e in enc], out=off[1:])
        blob = b"".join(enc)
    cap = rows * (cols * (26 if repr_floats else 7) + 40) + (len(blob) if blob else 0) + 64
    lib = native.load()
    for _ in range(2):
        buf = ctypes.create_string_buffer(cap)
        written = ctypes.c_size_t()
        rc = lib.sd_host_format_rows(3 if repr_floats else _KINDS[m.dtype], native.ptr(m), rows, cols, m.strides[0] // m.itemsize,
                                     blob, native.ptr(off), buf, cap, ctypes.byref(written), threads)
        if rc == native.SD_OK:
            return buf.raw[:written.value]
        if rc != native.SD_ERR_WORKSPACE:
            raise native.NativeCallError("sd_host_format_rows", rc, native.last_error())
        cap = written.value + 64
    raise native.NativeCallError("sd_host_format_rows", rc, native.last_error())


def write_matrix(path, header: str, names, matrix, chunk_rows=65536, threads=0, repr_floats=False):
    """header line + one formatted row per junction, streamed in row chunks."""
    with open(path, "wb") as out:
        out.write(header.encode())
        for r0 in range(0, matrix.shape[0], chunk_rows):
            out.write(format_rows(matrix[r0:r0 + chunk_rows], names[r0:r0 + chunk_rows], threads, repr_floats))


def read_table(path, threads=0):
    """(raw header line, row names, float64 [rows, cols] matrix) of a name<TAB>values table,
    parsed natively (sd_host_table_*).  ValueError on ragged rows / non-numeric fields."""
    import os
    lib = native.load()
    rows, cols = ctypes.c_int64(), ctypes.c_int32()
    hbytes, nbytes = ctypes.c_int64(), ctypes.c_int64()
    handle = lib.sd_host_table_open(os.fsencode(path), ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(hbytes),
                                    ctypes.byref(nbytes))
    if not handle:
        raise FileNotFoundError(native.last_error())
    handle = ctypes.c_void_p(handle)
    try:
        header = ctypes.create_string_buffer(max(hbytes.value, 1))
        names = ctypes.create_string_buffer(max(nbytes.value, 1))
        off = np.zeros(rows.value + 1, dtype=np.int64)
        values = np.empty((rows.value, cols.value), dtype=np.float64)
        rc = lib.sd_host_table_read(handle, header, names, native.ptr(off), native.ptr(values), max(cols.value, 1), threads)
        if rc != native.SD_OK:
            raise ValueError(native.last_error())
        blob = names.raw[:nbytes.value]
        o = off.tolist()
        return header.raw[:hbytes.value].decode(), [blob[o[i]:o[i + 1]].decode() for i in range(rows.value)], values
    finally:
        lib.sd_host_table_close(handle)
