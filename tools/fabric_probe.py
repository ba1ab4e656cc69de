#!/usr/bin/env python
"""Host <-> device link of the box, with no kernel in the way: pinned H2D alone, D2H alone and
both at once (the ceiling of every host-buffer pipeline), at several chunk sizes; then
sd_quant_ps_host at several row-block sizes next to it.  One process per GPU under torchrun
measures the aggregate.  Prints one JSON object per line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from splicedice_b200 import ops, synth  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    J, S = 400_000, 1000
    nbytes = J * S * 4
    for chunk_mb in (() if "--only-default" in sys.argv else (4, 32, 256)):
        r = bench.fabric_probe(dev, nbytes, nbytes, chunk_mb=chunk_mb, world=world)
        if rank == 0:
            print(json.dumps({"chunk_mb": chunk_mb, **r}), flush=True)
    if "--pipeline" in sys.argv:
        cl = ops.cluster_build(*synth.junction_arrays(J, 20261018)[:4])
        rp, ci = cl["row_ptr"].cpu().numpy(), cl["col_idx"].cpu().numpy()
        h_counts = torch.empty((J, S), dtype=torch.int32).pin_memory()
        h_counts.copy_(ops.synth_counts(1, 0, J, S, device=dev))
        h_ps = torch.empty((J, S), dtype=torch.float32).pin_memory()
        for mb in ((0,) if "--only-default" in sys.argv else (0, 8, 32, 128)):
            if mb:
                os.environ["SD_QUANT_HOST_BLOCK_MB"] = str(mb)
            elif "--only-default" not in sys.argv:
                os.environ.pop("SD_QUANT_HOST_BLOCK_MB", None)
            ops.quant_ps_host(h_counts, rp, ci, out=h_ps, device=local)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                ops.quant_ps_host(h_counts, rp, ci, out=h_ps, device=local)
                ts.append(time.perf_counter() - t0)
            if rank == 0:
                print(json.dumps({"sd_quant_ps_host_block_mb": mb or "default", "ms": [round(t * 1e3, 2) for t in ts]}), flush=True)


if __name__ == "__main__":
    main()
