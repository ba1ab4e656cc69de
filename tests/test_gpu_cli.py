"""The host commands end to end on the GPU against files written by the reference itself
(tests/golden/*/expected, made by oracle/gen_golden.py): byte-for-byte for every quant /
counts_to_ps / ir_table file, 1e-9 relative for the pairwise p-values."""
import argparse
import filecmp
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

QUANT_CASES = ["survey_vector", "cli_8x", "cli_8x_lownan", "cli_5x_pairwise"]


def _args(module, argv):
    p = argparse.ArgumentParser()
    module.add_parser(p)
    return p.parse_args(argv)


def _manifest(case_dir, tmp_path):
    out = tmp_path / "manifest.txt"
    with open(os.path.join(case_dir, "input", "manifest.txt")) as src, open(out, "w") as dst:
        for line in src:
            row = line.rstrip("\n").split("\t")
            row[1] = os.path.join(case_dir, row[1])
            dst.write("\t".join(row) + "\n")
    return str(out)


def _same(a, b):
    assert filecmp.cmp(a, b, shallow=False), f"{a} differs from {b}"


@pytest.mark.parametrize("case", QUANT_CASES)
def test_quant_files_are_byte_identical(case, golden_dir, tmp_path, capsys):
    from splicedice_b200 import counts_to_ps, quant
    case_dir = os.path.join(golden_dir, case)
    over = json.load(open(os.path.join(case_dir, "quant_args.json")))
    argv = ["-m", _manifest(case_dir, tmp_path), "-o", str(tmp_path / "out")]
    if over.get("lowCoverageNan"):
        argv.append("--lowCoverageNan")
    if "minUnique" in over:
        argv += ["--minUnique", str(over["minUnique"])]
    if over.get("drim"):
        argv.append("--drim")
    quant.run_with(_args(quant, argv))
    exp = os.path.join(case_dir, "expected")
    suffixes = ("_allClusters.tsv", "_junctions.bed", "_inclusionCounts.tsv", "_allPS.tsv")
    for suffix in suffixes + (("_drimTable.tsv",) if over.get("drim") else ()):
        _same(str(tmp_path / f"out{suffix}"), os.path.join(exp, f"ref{suffix}"))
    assert "All done" in capsys.readouterr().out

    # counts_to_ps from the reference's own intermediate files, both entry points
    counts_tsv, clusters_tsv = os.path.join(exp, "ref_inclusionCounts.tsv"), os.path.join(exp, "ref_allClusters.tsv")
    counts_to_ps.run_with(_args(counts_to_ps, ["-i", counts_tsv, "-c", clusters_tsv, "-o", str(tmp_path / "c")]))
    _same(str(tmp_path / "c_allPS.tsv"), os.path.join(exp, "c2ps_c_allPS.tsv"))
    counts_to_ps.run_with(_args(counts_to_ps, ["-i", counts_tsv, "-r", "-o", str(tmp_path / "r")]))
    _same(str(tmp_path / "r_allPS.tsv"), os.path.join(exp, "c2ps_r_allPS.tsv"))
    _same(str(tmp_path / "r_allClusters.tsv"), os.path.join(exp, "c2ps_r_allClusters.tsv"))


@pytest.mark.parametrize("native_io", [True, False], ids=["native", "python"])
@pytest.mark.parametrize("variant", ["default", "strict", "lengths"])
def test_quant_mixed_sample_formats_byte_identical(variant, native_io, golden_dir, tmp_path):
    """SJ.out.tab + tagged BED + plain BED + leafcutter (+ .bam / unknown suffix) in one manifest,
    every filter on its boundary, three flag sets: the four files the reference wrote
    (SPLICEDICE.py:162-228 junction union, :257-295 counts and low cells, :297-310 PS)."""
    from splicedice_b200 import quant
    from tests.test_ingest_golden import absolute_manifest, variant_argv
    case_dir = os.path.join(golden_dir, "mixed_formats")
    over = json.load(open(os.path.join(case_dir, "variants.json")))[variant]
    argv = ["-m", absolute_manifest(case_dir, tmp_path), "-o", str(tmp_path / "out")] + variant_argv(over)
    if not native_io:
        argv.append("--pythonIO")
    quant.run_with(_args(quant, argv))
    for suffix in ("_allClusters.tsv", "_junctions.bed", "_inclusionCounts.tsv", "_allPS.tsv"):
        _same(str(tmp_path / f"out{suffix}"), os.path.join(case_dir, f"expected_{variant}", f"ref{suffix}"))


def test_quant_api_shapes_match_the_reference(golden_dir, tmp_path):
    """getClusters() -> dict of tuples with the reference's list order; calculatePsi() -> float32."""
    from splicedice_b200 import quant
    case_dir = os.path.join(golden_dir, "cli_8x")
    args = _args(quant, ["-m", _manifest(case_dir, tmp_path), "-o", str(tmp_path / "o")])
    job = quant.SPLICEDICE(args.manifest, args.output_prefix, args, run=False)
    job.manifest = job.parseManifest()
    job.junctions = job.getAllJunctions()
    clusters = job.getClusters()
    assert set(clusters) == job.junctions
    assert list(job.junctionIndex) == sorted(job.junctions)
    name = lambda j: f"{j[0]}:{j[1]}-{j[2]}:{j[3]}"  # noqa: E731
    for line in open(os.path.join(case_dir, "expected", "ref_allClusters.tsv")):
        key, _, members = line.rstrip("\n").partition("\t")
        chrom, coords, strand = key.split(":")
        left, right = coords.split("-")
        got = [name(j) for j in clusters[(chrom, int(left), int(right), strand)]]
        assert got == ([m for m in members.split(",")] if members else [])
    job.counts, job.low = job.getJunctionCounts()
    psi = job.calculatePsi()
    assert psi.dtype == np.float32 and psi.shape == (len(job.junctions), len(job.manifest))


@pytest.mark.parametrize("case", ["survey_vector", "cli_5x_pairwise"])
@pytest.mark.parametrize("mode", ["none", "pairwise", "all"])
def test_pairwise_files(case, mode, golden_dir, tmp_path):
    from splicedice_b200 import pairwise_fisher
    exp = os.path.join(golden_dir, case, "expected")
    out = str(tmp_path / "pw.tsv")
    pairwise_fisher.run_with(_args(pairwise_fisher, [
        "--inclusionSPLICEDICE", os.path.join(exp, "ref_inclusionCounts.tsv"),
        "-c", os.path.join(exp, "ref_allClusters.tsv"), "--multiple_test_correction", mode, "-o", out]))
    got = [l.rstrip("\n").split("\t") for l in open(out)]
    want = [l.rstrip("\n").split("\t") for l in open(os.path.join(exp, f"pairwise_{mode}.tsv"))]
    assert got[0] == want[0] and [r[0] for r in got] == [r[0] for r in want]
    g = np.array([[float(x) for x in r[1:]] for r in got[1:]])
    w = np.array([[float(x) for x in r[1:]] for r in want[1:]])
    np.testing.assert_allclose(g, w, rtol=1e-9, atol=0)
    assert np.array_equal(g == 1.0, w == 1.0)


def test_pairwise_filter_list_quirk(golden_dir, tmp_path):
    """With --filter_list only listed events are loaded, so exclusions sum over listed partners only
    (pairwise_fisher.py:53-56,158)."""
    from oracle import fisher_c, oracle_np
    from splicedice_b200 import pairwise_fisher
    exp = os.path.join(golden_dir, "cli_5x_pairwise", "expected")
    rows = [l.rstrip("\n").split("\t") for l in open(os.path.join(exp, "ref_inclusionCounts.tsv"))][1:]
    keep = [r[0] for r in rows[::2]]
    flt = tmp_path / "filter.txt"
    flt.write_text("\n".join(keep) + "\n")
    out = str(tmp_path / "pw.tsv")
    pairwise_fisher.run_with(_args(pairwise_fisher, [
        "--inclusionSPLICEDICE", os.path.join(exp, "ref_inclusionCounts.tsv"), "-c",
        os.path.join(exp, "ref_allClusters.tsv"), "--multiple_test_correction", "none", "-f", str(flt), "-o", out]))
    got = [l.rstrip("\n").split("\t") for l in open(out)][1:]
    assert [r[0] for r in got] == keep
    counts = np.array([[int(x) for x in r[1:]] for r in rows[::2]], dtype=np.int64)
    adj = {}
    for l in open(os.path.join(exp, "ref_allClusters.tsv")):
        f = l.split()
        adj[f[0]] = f[1].split(",") if len(f) > 1 else []
    idx = {n: i for i, n in enumerate(keep)}
    exc = np.array([counts[[idx[o] for o in adj[n] if o in idx]].sum(axis=0) if any(o in idx for o in adj[n])
                    else np.zeros(counts.shape[1], int) for n in keep])
    pa, pb = oracle_np.all_pairs(counts.shape[1])
    want = fisher_c.pairwise(counts, exc, pa, pb)
    np.testing.assert_allclose(np.array([[float(x) for x in r[1:]] for r in got]), want, rtol=1e-9, atol=0)


def test_ir_table_files(golden_dir, tmp_path):
    from splicedice_b200 import ir_table
    d = os.path.join(golden_dir, "ir_small")
    samples = json.load(open(os.path.join(d, "samples.json")))
    counts = ir_table.getInclusionCounts(os.path.join(d, "counts.tsv"))
    clusters = ir_table.getClusters(os.path.join(d, "clusters.tsv"))
    args = argparse.Namespace(allJunctions=True, makeRSDtable=True, singleJunctionCalculation=False, RSDthreshold=1.0)
    kept, IR, RSD = ir_table.calculateIR(samples, os.path.join(d, "cov"), counts, clusters, None, args)
    ir_table.writeIRtable(samples, str(tmp_path / "a"), kept, IR)
    ir_table.writeRSDtable(samples, str(tmp_path / "a"), kept, RSD)
    _same(str(tmp_path / "a_intron_retention.tsv"), os.path.join(d, "expected", "ref_intron_retention.tsv"))
    _same(str(tmp_path / "a_intron_retention_RSD.tsv"), os.path.join(d, "expected", "ref_intron_retention_RSD.tsv"))
    args.singleJunctionCalculation = True
    kept, IR, RSD = ir_table.calculateIR(samples, os.path.join(d, "cov"), counts, None, None, args)
    ir_table.writeIRtable(samples, str(tmp_path / "s"), kept, IR)
    _same(str(tmp_path / "s_intron_retention.tsv"), os.path.join(d, "expected", "ref_single_intron_retention.tsv"))


def test_cli_dispatcher(golden_dir, tmp_path):
    from splicedice_b200 import __main__ as cli
    exp = os.path.join(golden_dir, "survey_vector", "expected")
    cli.main(["counts_to_ps", "-i", os.path.join(exp, "ref_inclusionCounts.tsv"), "-c",
              os.path.join(exp, "ref_allClusters.tsv"), "-o", str(tmp_path / "x")])
    _same(str(tmp_path / "x_allPS.tsv"), os.path.join(exp, "c2ps_c_allPS.tsv"))


def test_degenerate_inputs(tmp_path):
    """No junction survives the filters / a single sample: the files the reference would write
    (headers only; a trailing tab where the value list is empty)."""
    from splicedice_b200 import pairwise_fisher, quant
    bed = tmp_path / "a.junc.bed"
    bed.write_text("chr1\t100\t120\tj\t50\t+\n")                       # too short: filtered out
    man = tmp_path / "m.txt"
    man.write_text(f"s0\t{bed}\tm\tc\ns1\t{bed}\tm\tc\n")
    quant.run_with(_args(quant, ["-m", str(man), "-o", str(tmp_path / "e")]))
    assert open(tmp_path / "e_allClusters.tsv").read() == ""
    assert open(tmp_path / "e_junctions.bed").read() == ""
    assert open(tmp_path / "e_inclusionCounts.tsv").read() == "cluster\ts0\ts1\n"
    assert open(tmp_path / "e_allPS.tsv").read() == "cluster\ts0\ts1\n"
    # one sample -> no pairs: the reference writes "clusterID\t" and "event\t" lines
    counts = tmp_path / "c.tsv"
    counts.write_text("cluster\ts0\nchr1:1-100:+\t5\nchr1:50-200:+\t7\n")
    clusters = tmp_path / "l.tsv"
    clusters.write_text("chr1:1-100:+\tchr1:50-200:+\nchr1:50-200:+\tchr1:1-100:+\n")
    out = tmp_path / "p.tsv"
    pairwise_fisher.run_with(_args(pairwise_fisher, ["--inclusionSPLICEDICE", str(counts), "-c", str(clusters),
                                                     "-o", str(out)]))
    assert out.read_text() == "clusterID\t\nchr1:1-100:+\t\nchr1:50-200:+\t\n"
