#!/usr/bin/env python
"""First-use costs on a GPU box: torch import, CUDA context, library load, first call of each
kernel family (B200, this image: import torch 2.8 s, context 1.25 s, first cluster build 0.05 s;
the CLI's first device stage pays the first two once)."""
import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t=time.perf_counter()
import torch
print(f"import torch {time.perf_counter()-t:.2f}s"); t=time.perf_counter()
from splicedice_b200 import native, ops, synth
print(f"import package {time.perf_counter()-t:.2f}s"); t=time.perf_counter()
native.load()
print(f"native.load {time.perf_counter()-t:.2f}s"); t=time.perf_counter()
torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
print(f"cuda init {time.perf_counter()-t:.2f}s"); t=time.perf_counter()
arrays = synth.junction_arrays(200000, 11)[:4]
print(f"synth {time.perf_counter()-t:.2f}s"); t=time.perf_counter()
cl = ops.cluster_build(*arrays); torch.cuda.synchronize()
print(f"first cluster_build {time.perf_counter()-t:.2f}s"); t=time.perf_counter()
cl = ops.cluster_build(*arrays); torch.cuda.synchronize()
print(f"second cluster_build {time.perf_counter()-t:.3f}s"); t=time.perf_counter()
c = ops.synth_counts(1, 0, 200000, 64, device=0); torch.cuda.synchronize()
print(f"first synth_counts {time.perf_counter()-t:.3f}s"); t=time.perf_counter()
r = ops.quant_ps(c, cl["row_ptr"], cl["col_idx"]); torch.cuda.synchronize()
print(f"first quant_ps {time.perf_counter()-t:.3f}s"); t=time.perf_counter()
