"""The C-ABI library: builds, loads on a machine without a GPU, exports exactly what
include/splicedice_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "splicedice_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from splicedice_b200 import native
    assert declared_symbols() == sorted(native.EXPORTED_SYMBOLS)


def test_library_loads_and_exports_every_declared_symbol():
    from splicedice_b200 import native
    lib = native.load()
    raw = ctypes.CDLL(native.library_path())
    for name in declared_symbols():
        assert hasattr(raw, name), name
    assert lib.sd_version() == native.ABI_VERSION
    assert isinstance(native.last_error(), str)


def test_argument_errors_are_reported_without_a_device():
    from splicedice_b200 import native
    native.load()
    with pytest.raises(native.NativeCallError) as e:
        native.call("sd_quant_ps", -1, 4, None, 4, None, None, None, 0, None, 0, None, 0, None, 0, 0, 0, 0, None)
    assert e.value.code == native.SD_ERR_INVALID and "negative" in str(e.value)
    with pytest.raises(native.NativeCallError):
        native.call("sd_fisher_tables", 3, None, None, None, None, None, None)


def test_no_cpu_fallback():
    """Without CUDA the operators refuse to run instead of computing on the host."""
    import torch
    from splicedice_b200 import native, ops
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(native.NativeLibraryError):
        ops.require_cuda()
    import numpy as np
    with pytest.raises(native.NativeLibraryError):
        ops.fisher_tables(np.array([1]), np.array([2]), np.array([3]), np.array([4]))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "splicedice_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text or f.endswith(".md"), f
