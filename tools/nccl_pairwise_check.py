#!/usr/bin/env python
"""Multi-GPU check of distributed.pairwise_sharded: every rank runs exclusion sums + Fisher on its
row slab, the per-pair Benjamini-Hochberg correction exchanges row slabs for column blocks over
NCCL (all-to-all), and every rank compares its adjusted rows with the single-GPU result bit for
bit.  Launch with torchrun, one rank per GPU:  ... tools/nccl_pairwise_check.py [junctions] [samples]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from splicedice_b200 import distributed, ops, synth  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
J = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cl = ops.cluster_build(*synth.junction_arrays(J, 9)[:4])
rp, ci = cl["row_ptr"].cpu().numpy(), cl["col_idx"].cpu().numpy()
counts = synth.counts_host(10, 0, J, S) + synth.counts_host(11, 0, J, S)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize(); dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        out = fn()
        ev[i + 1].record()
    torch.cuda.synchronize(); dist.barrier()
    ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, float(t.item())


(p_adj, (r0, r1)), ms_all = timed(lambda: distributed.pairwise_sharded(counts, rp, ci, "pairwise"))
(p_raw, _), ms_none = timed(lambda: distributed.pairwise_sharded(counts, rp, ci, "none"))
parts = distributed.sharding.partition_rows(rp, ci, world)
_, ms_bh = timed(lambda: distributed.bh_columns_sharded(p_raw, parts, lambda c: ops.bh_adjust(c, "pairwise", out=c)))
_, ms_x = timed(lambda: distributed.columns_to_rows(distributed.rows_to_columns(p_raw, parts), parts, p_raw.shape[1]))

# single-GPU reference on this rank's device (the whole problem)
inc = torch.from_numpy(counts.astype(np.int32)).to(dev)
exc = ops.quant_ps(inc, rp, ci, want_f32=False, want_exc=True)["exc"]
pa, pb = ops.all_pairs(S)
single = ops.fisher_pairwise(inc, exc, pa, pb)
same_raw = torch.equal(p_raw.view(torch.int64), single[r0:r1].view(torch.int64))
ops.bh_adjust(single, "pairwise", out=single)
same_adj = torch.equal(p_adj.view(torch.int64), single[r0:r1].view(torch.int64))
flag = torch.tensor([int(same_raw), int(same_adj)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    P = S * (S - 1) // 2
    print(json.dumps({"world": world, "junctions": J, "samples": S, "pairs": P,
                      "pairwise_with_bh_ms": ms_all, "pairwise_no_correction_ms": ms_none,
                      "bh_exchange_plus_adjust_ms": ms_bh, "all_to_all_round_trip_ms": ms_x,
                      "all_to_all_bytes_per_rank_each_way": int((r1 - r0) * P * 8 * (world - 1) / world),
                      "raw_bit_identical_to_single_gpu": bool(flag[0].item()),
                      "adjusted_bit_identical_to_single_gpu": bool(flag[1].item()),
                      "note": "times include the host-side slab upload of pairwise_sharded; max over ranks"}))
dist.destroy_process_group()
sys.exit(0 if flag.min().item() else 1)
