#!/usr/bin/env python
"""sd_quant_ps_host with pinned buffers at several row-block sizes (SD_QUANT_HOST_BLOCK_MB)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from splicedice_b200 import ops, synth  # noqa: E402

J, S = 400_000, 1000
dev = torch.device("cuda", 0)
cl = ops.cluster_build(*synth.junction_arrays(J, 20261018)[:4])
rp, ci = cl["row_ptr"].cpu().numpy(), cl["col_idx"].cpu().numpy()
h_counts = torch.empty((J, S), dtype=torch.int32).pin_memory()
h_counts.copy_(ops.synth_counts(1, 0, J, S, device=dev))
h_ps = torch.empty((J, S), dtype=torch.float32).pin_memory()
# plain copies as the PCIe reference
d = torch.empty((J, S), dtype=torch.int32, device=dev)
for name, fn in (("H2D 1.6 GB", lambda: d.copy_(h_counts, non_blocking=True)),
                 ("D2H 1.6 GB", lambda: h_counts.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: {dt * 1e3:.1f} ms = {1.6 / dt:.1f} GB/s")
for mb in (8, 16, 32, 64, 128, 256):
    os.environ["SD_QUANT_HOST_BLOCK_MB"] = str(mb)
    ops.quant_ps_host(h_counts, rp, ci, out=h_ps)
    t0 = time.perf_counter()
    for _ in range(3):
        ops.quant_ps_host(h_counts, rp, ci, out=h_ps)
    dt = (time.perf_counter() - t0) / 3
    print(f"block {mb:4d} MB: {dt * 1e3:.1f} ms/call = {J * S / dt:.3e} cells/s")
