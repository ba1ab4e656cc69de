import os, sys, time, torch, numpy as np
sys.path.insert(0, ".")
from splicedice_b200 import ops, synth
J, S = 200_000, 64
dev = torch.device("cuda", 0)
cl = ops.cluster_build(*synth.junction_arrays(J, 20261025)[:4])
rp, ci = cl["row_ptr"].cpu().numpy(), cl["col_idx"].cpu().numpy()
inc = ops.synth_counts(8, 0, J, S, device=dev) + ops.synth_counts(9, 0, J, S, device=dev)
h_inc = torch.empty(inc.shape, dtype=torch.int32).pin_memory(); h_inc.copy_(inc)
pa, pb = ops.all_pairs(S)
out = torch.empty((J, len(pa)), dtype=torch.float64).pin_memory()
for mb in (64, 32, 128, 16):
    os.environ["SD_FISHER_HOST_BLOCK_MB"] = str(mb)
    ops.pairwise_host(h_inc, rp, ci, pa, pb, out=out)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); ops.pairwise_host(h_inc, rp, ci, pa, pb, out=out); ts.append((time.perf_counter() - t0) * 1e3)
    print("block_mb", mb, [round(t, 2) for t in ts], flush=True)
# plain D2H of the same bytes
d = torch.empty((J, len(pa)), dtype=torch.float64, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter(); out.copy_(d, non_blocking=True); torch.cuda.synchronize(); print("plain D2H ms", (time.perf_counter() - t0) * 1e3)
