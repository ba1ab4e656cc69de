"""Install the UNMODIFIED reference into ``oracle/_ref/`` (TEST INFRASTRUCTURE / CPU BASELINE).

The reference is a pure-python package, so "building" it is a pip install of its own source tree
(``python -m pip install --no-index --no-build-isolation --no-deps --target oracle/_ref`` from a
copy under /tmp, because ``/root/reference`` is read-only and setuptools writes ``build/`` and
``*.egg-info`` next to ``setup.py``).  ``oracle/_ref/`` is git-ignored (no reference source ever
enters the history) but NOT gpurun-ignored, so the installed package travels to the GPU box with
the tree exactly like the in-tree ``.so`` files, and ``bench.py``'s CPU legs time the reference's
own ``SPLICEDICE.getClusters`` / ``calculatePsi`` / ``pairwise_fisher`` loop there
(``cpu_baseline.kind = "reference"``).  Runs only where ``/root/reference`` exists (the build
container); elsewhere the already-installed copy is used as is.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SPLICEDICE_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
MARKER = os.path.join(OUT_DIR, "splicedice", "SPLICEDICE.py")


def installed() -> bool:
    return os.path.isfile(MARKER)


def source_present() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "splicedice", "SPLICEDICE.py"))


def _stale() -> bool:
    if not installed():
        return True
    src = os.path.join(REF_SRC, "splicedice")
    for name in os.listdir(src):
        if name.endswith(".py"):
            dst = os.path.join(OUT_DIR, "splicedice", name)
            if not os.path.isfile(dst) or os.path.getmtime(os.path.join(src, name)) > os.path.getmtime(dst):
                return True
    return False


def build(force: bool = False) -> str | None:
    """Returns the install directory, or None when neither the source nor an installed copy exists."""
    if not source_present():
        return OUT_DIR if installed() else None
    if not force and not _stale():
        return OUT_DIR
    if os.path.isdir(OUT_DIR):
        shutil.rmtree(OUT_DIR)
    with tempfile.TemporaryDirectory(prefix="splicedice_ref_") as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, work, ignore=shutil.ignore_patterns(".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--disable-pip-version-check", "--find-links", "/opt/wheelhouse", "--target", OUT_DIR, work]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    if not installed():
        raise RuntimeError(f"pip reported success but {MARKER} is missing")
    return OUT_DIR


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
