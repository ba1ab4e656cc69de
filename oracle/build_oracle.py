"""Build the C part of the oracle (TEST INFRASTRUCTURE): oracle/_build/libfisher_oracle.so."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libfisher_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "fisher_oracle.c")
    if (not force and os.path.isfile(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(src)):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-o", LIB, src, "-lquadmath", "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
