#!/usr/bin/env python
"""End-to-end `quant` CLI on synthetic .junc.bed files: stage timings with the native readers /
writers against the plain python readers (the reference's way of parsing).

    python tools/e2e_quant_cli.py [samples] [junctions] [workdir]
"""
import argparse
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from splicedice_b200 import quant, synth  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    J = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
    work = sys.argv[3] if len(sys.argv) > 3 else tempfile.mkdtemp(prefix="sd_e2e_")
    os.makedirs(work, exist_ok=True)
    js = synth.junction_tuples(J, 11)
    counts = synth.counts_host(12, 0, J, S)
    t0 = time.time()
    man = os.path.join(work, "manifest.txt")
    with open(man, "w") as m:
        for s in range(S):
            path = os.path.join(work, f"s{s}.junc.bed")
            col = counts[:, s]
            with open(path, "w") as f:
                f.write("".join(f"{c}\t{l}\t{r}\tj\t{v}\t{st}\n" for (c, l, r, st), v in zip(js, col.tolist()) if v))
            m.write(f"s{s}\t{path}\tmeta\tcond{s % 2}\n")
    print(f"wrote {S} sample files x {J} junctions in {time.time() - t0:.1f} s")
    outs = {}
    for mode, extra in (("native", []), ("python", ["--pythonIO"])):
        p = argparse.ArgumentParser()
        quant.add_parser(p)
        args = p.parse_args(["-m", man, "-o", os.path.join(work, f"out_{mode}"), *extra])
        buf = io.StringIO()
        t0 = time.time()
        with contextlib.redirect_stdout(buf):
            quant.run_with(args)
        total = time.time() - t0
        stages = [l.strip() for l in buf.getvalue().splitlines()]
        print(f"--- {mode} I/O: {total:.2f} s total")
        for a, b in zip(stages[::2], stages[1::2]):
            print(f"    {a:45s} {b}")
        outs[mode] = total
    same = all(open(os.path.join(work, f"out_native{sfx}"), "rb").read() == open(os.path.join(work, f"out_python{sfx}"), "rb").read()
               for sfx in ("_allClusters.tsv", "_junctions.bed", "_inclusionCounts.tsv", "_allPS.tsv"))
    print("outputs identical:", same, f"| speed-up {outs['python'] / outs['native']:.1f}x")


if __name__ == "__main__":
    main()
