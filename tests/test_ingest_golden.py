"""Sample-file ingest against the reference itself: a manifest mixing
SJ.out.tab, tagged ``splicedicebed``, plain BED and leafcutter files (plus a .bam and an unknown
suffix) whose lines sit on every filter boundary, run through the UNMODIFIED reference by
oracle/gen_golden.py (tests/golden/mixed_formats) under three flag sets.

Reference rules pinned here: SPLICEDICE.py:162-181 (SJ: strict length bounds, strand map, motif
set, unique (+ multi) score), :183-208 (tagged BED: filters only for ``a:?``), :210-226 (plain BED /
leafcutter: inclusive bounds), :257-295 (counts: last line wins, no per-sample filter, low cells).
The native reader (csrc/sd_ingest.cpp) is host code inside the product library, so these run
without a GPU."""
import json
import os

import numpy as np
import pytest

from tests.test_host_logic import _quant_job

CASE = "mixed_formats"
VARIANTS = ["default", "strict", "lengths"]


def variant_argv(over):
    argv = []
    for key in ("noMultimap", "lowCoverageNan", "drim"):
        if over.get(key):
            argv.append(f"--{key}")
    for key in ("minOverhang", "minEntropy", "minUnique", "maxLength", "minLength"):
        if key in over:
            argv += [f"--{key}", str(over[key])]
    return argv


def absolute_manifest(case_dir, tmp_path):
    out = tmp_path / "manifest.txt"
    with open(os.path.join(case_dir, "input", "manifest.txt")) as src, open(out, "w") as dst:
        for line in src:
            row = line.rstrip("\n").split("\t")
            row[1] = os.path.join(case_dir, row[1])
            dst.write("\t".join(row) + "\n")
    return str(out)


def _name_to_tuple(name):
    chrom, coords, strand = name.rsplit(":", 2)
    left, right = coords.split("-")
    return (chrom, int(left), int(right), strand)


@pytest.mark.parametrize("native_io", [True, False], ids=["native", "python"])
@pytest.mark.parametrize("variant", VARIANTS)
def test_ingest_reproduces_the_reference_files(variant, native_io, golden_dir, tmp_path):
    case_dir = os.path.join(golden_dir, CASE)
    over = json.load(open(os.path.join(case_dir, "variants.json")))[variant]
    exp = os.path.join(case_dir, f"expected_{variant}")
    job = _quant_job(absolute_manifest(case_dir, tmp_path), tmp_path / "o", variant_argv(over), native_io=native_io)
    assert [s.type for s in job.manifest] == ["SJ", "SJ", "splicedicebed", "splicedicebed", "bed", "leafcutter", "bam", "unknown"]
    job.junctions = job.getAllJunctions()
    want_rows = [_name_to_tuple(l.split("\t")[3]) for l in open(os.path.join(exp, "ref_junctions.bed"))]
    assert sorted(job.junctions) == want_rows and len(want_rows) > 150
    job._rows = want_rows
    job.junctionIndex = {j: r for r, j in enumerate(want_rows)}
    counts, _low = job.getJunctionCounts()
    lines = open(os.path.join(exp, "ref_inclusionCounts.tsv")).read().splitlines()
    assert lines[0].split("\t")[1:] == [s.name for s in job.manifest]
    want = np.array([[int(x) for x in l.split("\t")[1:]] for l in lines[1:]], dtype=np.int64)
    np.testing.assert_array_equal(counts, want)
    assert (counts[:, 6:] == 0).all()                       # .bam and unknown suffix: zero columns


@pytest.mark.needs_reference
@pytest.mark.parametrize("variant", VARIANTS)
def test_ingest_equals_the_live_reference(variant, golden_dir, tmp_path):
    """Junction union, count matrix and the low-cell list against the reference's own methods
    executed here on the same files."""
    from oracle import ref_harness as rh
    case_dir = os.path.join(golden_dir, CASE)
    over = json.load(open(os.path.join(case_dir, "variants.json")))[variant]
    manifest = absolute_manifest(case_dir, tmp_path)
    mod = rh.load("SPLICEDICE")
    rh.reset_sample_state()
    ref = object.__new__(mod.SPLICEDICE)
    ref.args = rh.quant_args(**over)
    ref.manifestFilename = manifest
    ref.manifest = ref.parseManifest()
    ref.junctions = ref.getAllJunctions()
    ref.clusters = ref.getClusters()
    ref.junctionIndex = {j: i for i, j in enumerate(sorted(ref.clusters))}
    ref_counts, ref_low = ref.getJunctionCounts()
    for native_io in (True, False):
        job = _quant_job(manifest, tmp_path / "o", variant_argv(over), native_io=native_io)
        job.junctions = job.getAllJunctions()
        assert job.junctions == ref.junctions
        job._rows = sorted(job.junctions)
        job.junctionIndex = {j: r for r, j in enumerate(job._rows)}
        counts, low = job.getJunctionCounts()
        np.testing.assert_array_equal(counts, ref_counts.astype(np.int64))
        assert sorted(set(map(tuple, low))) == sorted(set(ref_low))
        if over.get("lowCoverageNan"):
            assert len(ref_low) > 20
