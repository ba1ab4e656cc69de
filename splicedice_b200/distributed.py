"""One process per GPU: row-slab sharding of the quant / pairwise work over torch.distributed.

The path has no exchange step (adjacency never leaves an overlap component, see sharding.py):
every rank computes its slab independently and results return by host gather.  The only
collective is the optional all-gather of PS row blocks (NCCL over NVLink when the slabs are CUDA
tensors and the backend is nccl; gloo on CPU tensors in the tests), for a consumer that wants the
whole matrix on every GPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import sharding


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def sharded_rows(counts, row_ptr, col_idx, slab_fn, gather=True, weights=None):
    """Run ``slab_fn(counts[r0:r1], local_row_ptr, local_col_idx) -> tensor[r1 - r0, S]`` on this
    rank's slab.  With ``gather`` every rank receives the full [J, S] tensor (slabs padded to a
    common height for the all-gather); otherwise returns (local tensor, (r0, r1))."""
    rank, world = _world()
    n_samples = counts.shape[1]
    if weights is None:
        weights = sharding.row_weights(row_ptr, n_samples)
    parts = sharding.partition_rows(row_ptr, col_idx, world, weights)
    r0, r1 = parts[rank]
    rp, ci = sharding.shard_csr(row_ptr, col_idx, r0, r1)
    local = slab_fn(counts[r0:r1], rp, ci)
    if not gather:
        return local, (r0, r1)
    if world == 1:
        return local
    height = max(b - a for a, b in parts)
    padded = torch.zeros((height,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: r1 - r0] = local
    blocks = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(blocks, padded)
    return torch.cat([blocks[k][: parts[k][1] - parts[k][0]] for k in range(world)])


def quant_ps_sharded(counts_host, row_ptr, col_idx, device=None, gather=True, low_mask=None):
    """PS of a host int32 count matrix with the rows split over the ranks' GPUs.  Rank r uploads
    only its slab.  Returns the full float32 matrix on every rank's GPU (``gather``) or the local
    slab and its row range."""
    from . import ops
    rank, _ = _world()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)

    def slab(counts_slab, rp, ci):
        c = torch.from_numpy(np.ascontiguousarray(counts_slab, dtype=np.int32)).to(dev)
        n = c.shape[1]
        if n % 4:                                  # 16-byte row alignment for the tiled kernels
            buf = torch.zeros((c.shape[0], (n + 3) // 4 * 4), dtype=torch.int32, device=dev)
            buf[:, :n] = c
            c = buf[:, :n]
        return ops.quant_ps(c, rp, ci, want_f32=True)["ps_f32"]

    if low_mask is not None:
        raise NotImplementedError("low_mask is applied per slab by the caller")
    return sharded_rows(counts_host, row_ptr, col_idx, slab, gather=gather)
