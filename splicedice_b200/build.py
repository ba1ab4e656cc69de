"""Build ``libsplicedice_b200.so`` (sm_100a only) in-tree with nvcc.

The shared library lives at ``splicedice_b200/_lib/libsplicedice_b200.so`` so that it travels
with the source tree (it is git-ignored, not gpurun-ignored).  ``python -m splicedice_b200.build``
rebuilds it; ``native.load()`` builds it on first use when nvcc is present and the sources are
newer than the library.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
LIB = os.path.join(OUT_DIR, "libsplicedice_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-O3",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    return None


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))


def stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in _deps())


def build_variant(tag: str, extra_flags) -> str:
    """An experiment build: every source recompiled with ``extra_flags`` (e.g. ``-DSD_FISHER_CUT_BITS=36``)
    into ``_lib/libsplicedice_b200.<tag>.so``; ``SPLICEDICE_B200_LIB=<that path>`` makes native.load() use it."""
    cc = nvcc()
    if cc is None:
        raise RuntimeError("nvcc not found")
    obj_dir = os.path.join(OUT_DIR, f"obj_{tag}")
    os.makedirs(obj_dir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
        r = subprocess.run([cc, *NVCC_FLAGS, *extra_flags, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources()))
    lib = os.path.join(OUT_DIR, f"libsplicedice_b200.{tag}.so")
    subprocess.run([cc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                    "-Xcompiler", "-fPIC", "-cudart", "static", "-lquadmath"], check=True)
    shutil.rmtree(obj_dir)
    return lib


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    if not force and not stale():
        return LIB
    cc = nvcc()
    if cc is None:
        raise RuntimeError("nvcc not found: cannot build libsplicedice_b200.so (no CPU fallback exists)")
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_time = max(os.path.getmtime(f) for f in _deps() if not f.endswith((".cu", ".cpp")))

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
        if (not force and os.path.isfile(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and os.path.getmtime(obj) > hdr_time):
            return obj
        cmd = [cc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if ptxas_info:
            cmd[1:1] = ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0 or ptxas_info:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [cc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-cudart", "static", "-lquadmath"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_info="--ptxas" in sys.argv))
