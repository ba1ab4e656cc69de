"""``--gpus N`` of the command-line tools (splicedice_b200/multigpu.py): one worker process per
GPU, row slabs at cluster boundaries, shared-memory host gather; pairwise joins a NCCL group for the
Benjamini-Hochberg column exchange / the gather of ``--multiple_test_correction all``.  Needs two
CUDA devices (skipped otherwise): the 2-GPU files are byte-identical to the reference-made goldens
and to the 1-GPU files, the 2-GPU matrices bit-identical to the 1-GPU ones."""
import argparse
import filecmp
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs two CUDA devices", allow_module_level=True)


def _args(module, argv):
    p = argparse.ArgumentParser()
    module.add_parser(p)
    return p.parse_args(argv)


def _same(a, b):
    assert filecmp.cmp(a, b, shallow=False), f"{a} differs from {b}"


@pytest.mark.parametrize("case,variant", [("cli_8x_lownan", None), ("mixed_formats", "strict"), ("mixed_formats", "default")])
def test_quant_two_gpus_files_are_byte_identical(case, variant, golden_dir, tmp_path):
    from splicedice_b200 import quant
    from tests.test_ingest_golden import absolute_manifest, variant_argv
    case_dir = os.path.join(golden_dir, case)
    if variant is None:
        over = json.load(open(os.path.join(case_dir, "quant_args.json")))
        exp = os.path.join(case_dir, "expected")
    else:
        over = json.load(open(os.path.join(case_dir, "variants.json")))[variant]
        exp = os.path.join(case_dir, f"expected_{variant}")
    argv = ["-m", absolute_manifest(case_dir, tmp_path), "-o", str(tmp_path / "out"), "--gpus", "2"] + variant_argv(over)
    quant.run_with(_args(quant, argv))
    for suffix in ("_allClusters.tsv", "_junctions.bed", "_inclusionCounts.tsv", "_allPS.tsv"):
        _same(str(tmp_path / f"out{suffix}"), os.path.join(exp, f"ref{suffix}"))


@pytest.mark.parametrize("mode", ["none", "pairwise", "all"])
def test_pairwise_two_gpus_file_equals_one_gpu(mode, golden_dir, tmp_path):
    from splicedice_b200 import pairwise_fisher
    exp = os.path.join(golden_dir, "cli_5x_pairwise", "expected")
    outs = []
    for gpus in (1, 2):
        out = str(tmp_path / f"pw{gpus}.tsv")
        pairwise_fisher.run_with(_args(pairwise_fisher, [
            "--inclusionSPLICEDICE", os.path.join(exp, "ref_inclusionCounts.tsv"), "-c",
            os.path.join(exp, "ref_allClusters.tsv"), "--multiple_test_correction", mode, "-o", out, "--gpus", str(gpus)]))
        outs.append(out)
    _same(*outs)


def test_two_gpu_matrices_are_bit_identical_on_a_larger_problem():
    """20,000 events x 10 samples (45 pairs) with the low-coverage mask / all three corrections."""
    from oracle import oracle_np
    from splicedice_b200 import multigpu, ops, pairwise_fisher
    from tests import util
    J, S = 20000, 10
    _, csr, counts = util.synthetic_problem(J, S, seed=5, zero_frac=0.15)
    low = (np.random.default_rng(2).random((J, S)) < 0.02).astype(np.uint8)
    one = ops.quant_ps_host(counts.astype(np.int32), csr["row_ptr"], csr["col_idx"], low_mask=low).numpy()
    two = multigpu.quant_ps(counts.astype(np.int32), csr["row_ptr"], csr["col_idx"], low_mask=low, n_gpus=2)
    np.testing.assert_array_equal(util.bits32(two), util.bits32(one))
    want = oracle_np.ps_f32(counts, csr["row_ptr"], csr["col_idx"], low_mask=low)
    np.testing.assert_array_equal(util.bits32(two), util.bits32(want))
    names = [f"chr1:{i}-{i + 1}:+" for i in range(J)]
    rp, ci = csr["row_ptr"], csr["col_idx"]
    clusters = {names[r]: [names[c] for c in ci[rp[r]:rp[r + 1]]] for r in range(J)}
    for mode in ("none", "pairwise", "all"):
        a = pairwise_fisher.pairwise_pvalues(names, counts.astype(np.float64), clusters, correction=mode, gpus=1)
        b = pairwise_fisher.pairwise_pvalues(names, counts.astype(np.float64), clusters, correction=mode, gpus=2)
        np.testing.assert_array_equal(util.bits64(a), util.bits64(b))


def test_gpus_argument_is_checked():
    from splicedice_b200 import multigpu
    with pytest.raises(RuntimeError):
        multigpu.check_gpus(torch.cuda.device_count() + 1)
    with pytest.raises(ValueError):
        multigpu.check_gpus(0)
