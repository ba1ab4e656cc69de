"""Row-slab sharding (host logic) and the 2-rank driver over gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle_np
from splicedice_b200 import sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _csr(n, seed):
    c, s, st, en, _, _ = synth.junction_arrays(n, seed)
    return oracle_np.cluster_csr(c, s, st, en)


def test_safe_cuts_match_brute_force():
    csr = _csr(600, 1)
    rp, ci = csr["row_ptr"], csr["col_idx"]
    safe = sharding.safe_cuts(rp, ci)
    J = len(rp) - 1
    rows = np.repeat(np.arange(J), np.diff(rp))
    for r in range(J + 1):
        crossing = np.any((rows < r) != (ci < r))
        assert safe[r] == (not crossing), r
    assert safe[0] and safe[J]


@pytest.mark.parametrize("n_shards", [1, 2, 3, 8])
def test_partition_is_closed_and_balanced(n_shards):
    csr = _csr(20000, 2)
    rp, ci = csr["row_ptr"], csr["col_idx"]
    w = sharding.row_weights(rp, 1000)
    parts = sharding.partition_rows(rp, ci, n_shards, w)
    assert parts[0][0] == 0 and parts[-1][1] == 20000
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    for r0, r1 in parts:
        lrp, lci = sharding.shard_csr(rp, ci, r0, r1)          # raises if an edge leaves the slab
        assert lrp[0] == 0 and lrp[-1] == len(lci)
    assert sharding.imbalance(parts, w) < 1.02
    # the shards' exclusion sums, stitched together, equal the unsharded ones
    counts = synth.counts_host(3, 0, 20000, 5)
    whole = oracle_np.exclusion_sums(counts, rp, ci)
    for r0, r1 in parts:
        lrp, lci = sharding.shard_csr(rp, ci, r0, r1)
        np.testing.assert_array_equal(oracle_np.exclusion_sums(counts[r0:r1], lrp, lci), whole[r0:r1])


def test_unsafe_cut_is_rejected():
    csr = _csr(500, 3)
    rp, ci = csr["row_ptr"], csr["col_idx"]
    bad = int(np.flatnonzero(~sharding.safe_cuts(rp, ci))[0])
    with pytest.raises(ValueError):
        sharding.shard_csr(rp, ci, 0, bad)


def test_more_shards_than_cuts():
    rp = np.array([0, 1, 2], dtype=np.int32)
    ci = np.array([1, 0], dtype=np.int32)                  # one 2-row component: no interior cut
    parts = sharding.partition_rows(rp, ci, 4)
    assert [b - a for a, b in parts].count(2) == 1 and sum(b - a for a, b in parts) == 2


def test_two_rank_gather_over_gloo(tmp_path):
    """world_size 2 on CPU (gloo): each rank takes its slab, the gathered matrix is complete and
    in output-row order.  The per-slab compute is injected (the oracle here; the CUDA operator in
    the product), so the test exercises partition + rebasing + gather, not arithmetic."""
    script = tmp_path / "worker.py"
    script.write_text(
        "import os, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import numpy as np, torch, torch.distributed as dist\n"
        "from oracle import oracle_np\n"
        "from splicedice_b200 import distributed, synth\n"
        "dist.init_process_group('gloo')\n"
        "rank = dist.get_rank()\n"
        "c, s, st, en, _, _ = synth.junction_arrays(5000, 7)\n"
        "csr = oracle_np.cluster_csr(c, s, st, en)\n"
        "counts = synth.counts_host(8, 0, 5000, 6)\n"
        "def slab_ps(counts_slab, rp, ci):\n"
        "    return torch.from_numpy(oracle_np.ps_f32(counts_slab, rp, ci))\n"
        "ps = distributed.sharded_rows(counts, csr['row_ptr'], csr['col_idx'], slab_ps, gather=True)\n"
        "want = oracle_np.ps_f32(counts, csr['row_ptr'], csr['col_idx'])\n"
        "assert np.array_equal(ps.numpy().view(np.uint32), want.view(np.uint32))\n"
        "dist.destroy_process_group()\n"
        "print('rank', rank, 'ok')\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_two_rank_bh_exchange_over_gloo(tmp_path):
    """The pairwise correction's exchange step on CPU tensors (gloo, world_size 2): row slabs ->
    column blocks -> per-column Benjamini-Hochberg (the numpy restatement stands in for
    sd_bh_adjust) -> back to row slabs; equals the correction of the whole matrix."""
    script = tmp_path / "worker_bh.py"
    script.write_text(
        "import os, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import numpy as np, torch, torch.distributed as dist\n"
        "from oracle import oracle_np\n"
        "from splicedice_b200 import distributed\n"
        "dist.init_process_group('gloo')\n"
        "rank, world = dist.get_rank(), dist.get_world_size()\n"
        "rng = np.random.default_rng(5)\n"
        "p = rng.random((1001, 7)) ** 3\n"
        "p[rng.random(p.shape) < 0.1] = 1.0\n"
        "parts = [(0, 430), (430, 1001)]\n"
        "def adjust(cols):\n"
        "    out = cols.numpy().copy()\n"
        "    for k in range(out.shape[1]):\n"
        "        out[:, k] = oracle_np.bh_adjust(out[:, k])\n"
        "    return torch.from_numpy(out)\n"
        "r0, r1 = parts[rank]\n"
        "cols = distributed.rows_to_columns(torch.from_numpy(p[r0:r1].copy()), parts)\n"
        "c0, c1 = distributed.column_blocks(7, world)[rank]\n"
        "assert np.array_equal(cols.numpy(), p[:, c0:c1])\n"
        "got = distributed.bh_columns_sharded(torch.from_numpy(p[r0:r1].copy()), parts, adjust)\n"
        "want = np.stack([oracle_np.bh_adjust(p[:, k]) for k in range(7)], axis=1)[r0:r1]\n"
        "assert np.array_equal(got.numpy().view(np.uint64), want.view(np.uint64))\n"
        "dist.destroy_process_group()\n"
        "print('rank', rank, 'ok')\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_column_blocks_cover_all_columns():
    from splicedice_b200 import distributed
    for n, w in [(7, 2), (2016, 8), (3, 8), (0, 2)]:
        blocks = distributed.column_blocks(n, w)
        assert len(blocks) == w and blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1


def test_numa_binding_is_a_harmless_hint():
    """bind_to_gpu_numa never raises: without NVML / a GPU it returns None and leaves the affinity alone."""
    from splicedice_b200 import distributed
    before = os.sched_getaffinity(0)
    got = distributed.bind_to_gpu_numa(0)
    assert got is None or set(got) <= before
    os.sched_setaffinity(0, before)
    os.environ["SD_NO_NUMA_BIND"] = "1"
    try:
        assert distributed.bind_to_gpu_numa(0) is None
    finally:
        del os.environ["SD_NO_NUMA_BIND"]
