#!/usr/bin/env python
"""One warm sd_cluster_build + sd_cluster_fill at J junctions (for an ncu launch list) and its wall time."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from splicedice_b200 import ops, synth
J = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
dev = torch.device("cuda", 0)
d = [torch.from_numpy(a).to(dev) for a in synth.junction_arrays(J, 3)[:4]]
ops.cluster_build(*d); torch.cuda.synchronize()
ts = []
for _ in range(7):
    t0 = time.perf_counter(); ops.cluster_build(*d); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print(f"cluster_build J={J}: ms per call {[round(t, 3) for t in ts]}")
