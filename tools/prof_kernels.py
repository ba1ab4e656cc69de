#!/usr/bin/env python
"""Small driver for ncu captures: runs the quant PS kernel (configs[1] shape, fewer rows) and/or
the pairwise Fisher kernel (configs[2] shape, fewer rows) a few times.

    python tools/prof_kernels.py quant [rows] [samples] [flags]
    python tools/prof_kernels.py fisher [rows] [samples]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from splicedice_b200 import ops, synth  # noqa: E402


def main():
    what = sys.argv[1]
    dev = torch.device("cuda", 0)
    if what == "quant":
        J = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
        S = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
        flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
        cl = ops.cluster_build(*synth.junction_arrays(J, 20261018)[:4])
        counts = ops.synth_counts(1, 0, J, S, device=dev)
        ps = torch.empty((J, S), dtype=torch.float32, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(4):
            if i == 1:
                e0.record()
            ops.quant_ps(counts, cl["row_ptr"], cl["col_idx"], out_f32=ps, flags=flags)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"quant {J}x{S} flags={flags}: {ms:.3f} ms/launch, {J * S * 8 / ms / 1e6:.0f} GB/s algorithmic")
    else:
        J = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
        S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
        cl = ops.cluster_build(*synth.junction_arrays(J, 20261025)[:4])
        inc = ops.synth_counts(8, 0, J, S, device=dev) + ops.synth_counts(9, 0, J, S, device=dev)
        exc = ops.quant_ps(inc, cl["row_ptr"], cl["col_idx"], want_f32=False, want_exc=True)["exc"]
        pa, pb = ops.all_pairs(S)
        out = torch.empty((J, len(pa)), dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3):
            if i == 1:
                e0.record()
            ops.fisher_pairwise(inc, exc, pa, pb, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        print(f"fisher {J}x{len(pa)}: {ms:.3f} ms/launch, {J * len(pa) / (ms * 1e-3):.3e} tests/s")


if __name__ == "__main__":
    main()
