#!/usr/bin/env python
"""sd_fisher_pairwise_host (host buffers in / out) at several p-value block sizes."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from splicedice_b200 import ops, synth  # noqa: E402

J, S = 200_000, 64
dev = torch.device("cuda", 0)
cl = ops.cluster_build(*synth.junction_arrays(J, 20261025)[:4])
inc = ops.synth_counts(8, 0, J, S, device=dev) + ops.synth_counts(9, 0, J, S, device=dev)
exc = ops.quant_ps(inc, cl["row_ptr"], cl["col_idx"], want_f32=False, want_exc=True)["exc"]
pa, pb = ops.all_pairs(S)
inc_h, exc_h = inc.cpu().numpy(), exc.cpu().numpy()
out = torch.empty((J, len(pa)), dtype=torch.float64).pin_memory()
for mb in (32, 64, 128, 256, 512):
    os.environ["SD_FISHER_HOST_BLOCK_MB"] = str(mb)
    ops.fisher_pairwise_host(inc_h, exc_h, pa, pb, out=out)
    t0 = time.perf_counter()
    for _ in range(2):
        ops.fisher_pairwise_host(inc_h, exc_h, pa, pb, out=out)
    dt = (time.perf_counter() - t0) / 2
    print(f"block {mb:4d} MB: {dt * 1e3:.1f} ms/call = {J * len(pa) / dt:.3e} tests/s")
