"""Row-text writers over ``sd_host_format_rows`` (native, multi-threaded, byte-identical to the
reference's per-cell f-strings)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import native

_KINDS = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 2}


def _prepare(matrix, names, repr_floats):
    m = np.asarray(matrix)
    if m.ndim != 2 or m.dtype not in _KINDS or (repr_floats and m.dtype != np.float64):
        raise TypeError("format_rows: 2-D float32 / float64 / int32 matrix expected")
    if m.strides[1] != m.itemsize:
        m = np.ascontiguousarray(m)
    rows, cols = m.shape
    blob = off = None
    if names is not None and rows:
        enc = [n.encode() for n in names]
        if len(enc) != rows:
            raise ValueError("format_rows: one name per row expected")
        off = np.zeros(rows + 1, dtype=np.int64)
        np.cumsum([len(e) for e in enc], out=off[1:])
        blob = b"".join(enc)
    return m, blob, off


def _cell_bytes(repr_floats):
    return 26 if repr_floats else 7


def _format_into(m, blob, off, threads, repr_floats, buf):
    """Format the rows of ``m`` into the uint8 array ``buf`` (grown when too small); returns
    (buf, bytes written).  No zero-fill and no copy of the text on the python side."""
    rows, cols = m.shape
    lib = native.load()
    written = ctypes.c_size_t()
    for _ in range(2):
        rc = lib.sd_host_format_rows(3 if repr_floats else _KINDS[m.dtype], native.ptr(m), rows, cols,
                                     m.strides[0] // m.itemsize, blob, native.ptr(off), native.ptr(buf), buf.size,
                                     ctypes.byref(written), threads)
        if rc == native.SD_OK:
            return buf, written.value
        if rc != native.SD_ERR_WORKSPACE:
            break
        buf = np.empty(written.value + 64, dtype=np.uint8)
    raise native.NativeCallError("sd_host_format_rows", rc, native.last_error())


def format_rows(matrix, names=None, threads=0, repr_floats=False) -> bytes:
    """b"name\\tv\\tv...\\n" for every row of a 2-D float32 / float64 ('.3f') or int32 matrix;
    ``repr_floats``: float64 as python's str(p) (the pairwise writer)."""
    m, blob, off = _prepare(matrix, names, repr_floats)
    rows, cols = m.shape
    if rows == 0:
        return b""
    cap = rows * (cols * _cell_bytes(repr_floats) + 40) + (len(blob) if blob else 0) + 64
    buf, n = _format_into(m, blob, off, threads, repr_floats, np.empty(cap, dtype=np.uint8))
    return buf[:n].tobytes()


def _format_segments(m, blob, off, threads, repr_floats, buf):
    """As _format_into without the compaction: returns (buf, [(offset, length), ...]) — every
    formatter thread's text where it was written, in row order (sd_host_format_rows_segments)."""
    rows, cols = m.shape
    lib = native.load()
    max_seg = 256
    seg_off = np.empty(max_seg, dtype=np.int64)
    seg_len = np.empty(max_seg, dtype=np.int64)
    n_seg, needed = ctypes.c_int32(), ctypes.c_size_t()
    for _ in range(2):
        rc = lib.sd_host_format_rows_segments(3 if repr_floats else _KINDS[m.dtype], native.ptr(m), rows, cols,
                                              m.strides[0] // m.itemsize, blob, native.ptr(off), native.ptr(buf),
                                              buf.size, native.ptr(seg_off), native.ptr(seg_len), max_seg,
                                              ctypes.byref(n_seg), ctypes.byref(needed), threads)
        if rc == native.SD_OK:
            return buf, list(zip(seg_off[:n_seg.value].tolist(), seg_len[:n_seg.value].tolist()))
        if rc != native.SD_ERR_WORKSPACE:
            break
        buf = np.empty(needed.value + 64, dtype=np.uint8)
    raise native.NativeCallError("sd_host_format_rows_segments", rc, native.last_error())


def write_matrix(path, header: str, names, matrix, chunk_rows=None, threads=0, repr_floats=False):
    """header line + one formatted row per junction, streamed in row chunks of ~64 MB of text
    through two alternating buffers: chunk k is written by a helper thread while chunk k+1 is
    being formatted (both release the GIL).  Each formatter thread's slice goes to the file from
    where it was written — no compaction, no python-side copy."""
    from concurrent.futures import ThreadPoolExecutor
    m = np.asarray(matrix)
    cols = m.shape[1] if m.ndim == 2 else 0
    if chunk_rows is None:
        chunk_rows = max(1, min(65536, (64 << 20) // max(1, cols * _cell_bytes(repr_floats) + 40)))
    bufs = [None, None]

    def flush(out, buf, segments):
        view = memoryview(buf)
        for o, n in segments:
            out.write(view[o:o + n])

    with open(path, "wb") as out, ThreadPoolExecutor(1) as writer:
        out.write(header.encode())
        pending = [None, None]
        for k, r0 in enumerate(range(0, m.shape[0], chunk_rows)):
            part, blob, off = _prepare(m[r0:r0 + chunk_rows], None if names is None else names[r0:r0 + chunk_rows],
                                       repr_floats)
            if part.shape[0] == 0:
                continue
            slot = k & 1
            if pending[slot] is not None:
                pending[slot].result()               # this buffer's previous text is on its way to the file
            cap = part.shape[0] * (cols * _cell_bytes(repr_floats) + 40) + (len(blob) if blob else 0) + 64
            if bufs[slot] is None or bufs[slot].size < cap:
                bufs[slot] = np.empty(cap, dtype=np.uint8)
            bufs[slot], segments = _format_segments(part, blob, off, threads, repr_floats, bufs[slot])
            pending[slot] = writer.submit(flush, out, bufs[slot], segments)   # one writer thread: chunks stay in order
        for f in pending:
            if f is not None:
                f.result()


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
