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
    import torch.distributed as dist
    if world > 1:
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
        modes = [("1", 16), ("1", 32), ("0", 16), ("0", 32), ("0", 128)] if "--modes" in sys.argv else [(None, 0)]
        for u16, mb in modes:
            if u16 is not None:
                os.environ["SD_QUANT_HOST_U16"] = u16
                os.environ["SD_QUANT_HOST_BLOCK_MB"] = str(mb)
            ops.quant_ps_host(h_counts, rp, ci, out=h_ps, device=local)
            ts = []
            for _ in range(8):
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                ops.quant_ps_host(h_counts, rp, ci, out=h_ps, device=local)
                dt = time.perf_counter() - t0
                if world > 1:
                    t = torch.tensor([dt], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dt = float(t.item())
                ts.append(dt)
            if rank == 0:
                print(json.dumps({"sd_quant_ps_host": {"u16": u16, "block_mb": mb or "default"}, "n_gpus": world,
                                  "ms_max_over_ranks": [round(t * 1e3, 2) for t in ts]}), flush=True)


if __name__ == "__main__":
    main()
