"""Generate tests/golden/* by EXECUTING the reference in the build container.

Test infrastructure.  Run here (``python -m oracle.gen_golden``) where
``/root/reference`` exists; the fixtures it writes are committed and replayed by
the test-suite on boxes where the reference does not exist.  The reference ships
no tests or vectors of its own, so these reference-run outputs are
the parity pin.  Versions that produced the committed fixtures are recorded in
``tests/golden/MANIFEST.json``.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from splicedice_b200 import synth  # noqa: E402


# SURVEY.md §4 known-answer vector (3 samples)
SURVEY_ROWS = [
    ("chr1", 100, 300, "+", (34, 32, 22)),
    ("chr1", 100, 300, "-", (1, 38, 0)),
    ("chr1", 100, 400, "+", (25, 25, 37)),
    ("chr1", 300, 500, "+", (20, 36, 11)),
    ("chr1", 301, 600, "+", (10, 20, 32)),
    ("chr1", 700, 900, "+", (12, 0, 26)),
    ("chr10", 50, 250, "+", (3, 29, 15)),
    ("chr2", 50, 250, "+", (0, 25, 34)),
    ("chr2", 60, 200, "+", (7, 21, 22)),
]


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def write_bed_inputs(dirname, junctions, counts, names, shuffle_seed=0):
    """One ``<name>.junc.bed`` per sample (zero-count lines omitted) + manifest.txt."""
    os.makedirs(dirname, exist_ok=True)
    rng = np.random.default_rng(shuffle_seed)
    man = []
    for s, name in enumerate(names):
        path = os.path.join(dirname, f"{name}.junc.bed")
        order = rng.permutation(len(junctions))
        with open(path, "w") as f:
            for i in order:
                c, l, r, st = junctions[i]
                if counts[i, s] == 0:
                    continue
                f.write(f"{c}\t{l}\t{r}\tj{i}\t{int(counts[i, s])}\t{st}\n")
        man.append(f"{name}\t{path}\tmeta{s}\tcond{s % 2}\n")
    mpath = os.path.join(dirname, "manifest.txt")
    with open(mpath, "w") as f:
        f.writelines(man)
    return mpath


def run_reference_quant(manifest, prefix, **over):
    mod = rh.load("SPLICEDICE")
    rh.reset_sample_state()
    args = rh.quant_args(manifest=manifest, output_prefix=prefix, **over)
    _quiet(mod.run_with, args)


def run_reference_pairwise(counts_tsv, clusters_tsv, out, correction, filter_list=None):
    mod = rh.load("pairwise_fisher")
    args = rh._Args(inclusionSPLICEDICE=counts_tsv, clusters=clusters_tsv, chi2=False,
                    multiple_test_correction=correction, filter_list=filter_list, output=out)
    _quiet(mod.run_with, args)


def run_reference_counts_to_ps(counts_tsv, clusters_tsv, prefix, recluster=False):
    mod = rh.load("counts_to_ps")
    args = rh._Args(inclusion_counts=counts_tsv, clusters=None if recluster else clusters_tsv,
                    recluster=recluster, output_prefix=prefix)
    _quiet(mod.run_with, args)


def case_dir(name):
    d = os.path.join(GOLD, name)
    if os.path.isdir(d):
        shutil.rmtree(d)
    os.makedirs(d)
    return d


def relpath_manifest(path, base):
    """Rewrite absolute sample paths in a manifest as paths relative to `base`."""
    out = []
    for line in open(path):
        row = line.rstrip("\n").split("\t")
        row[1] = os.path.relpath(row[1], base)
        out.append("\t".join(row) + "\n")
    with open(path, "w") as f:
        f.writelines(out)


def gen_cli_case(name, junctions, counts, quant_over=None, pairwise=True):
    """Full file-based pipeline: quant -> counts_to_ps (-c and -r) -> pairwise (3 modes)."""
    quant_over = quant_over or {}
    d = case_dir(name)
    inp = os.path.join(d, "input")
    names = [f"s{i}" for i in range(counts.shape[1])]
    manifest = write_bed_inputs(inp, junctions, counts, names)
    exp = os.path.join(d, "expected")
    os.makedirs(exp)
    prefix = os.path.join(exp, "ref")
    run_reference_quant(manifest, prefix, **quant_over)
    run_reference_counts_to_ps(prefix + "_inclusionCounts.tsv", prefix + "_allClusters.tsv",
                               os.path.join(exp, "c2ps_c"))
    run_reference_counts_to_ps(prefix + "_inclusionCounts.tsv", None,
                               os.path.join(exp, "c2ps_r"), recluster=True)
    if pairwise:
        for mode in ("none", "pairwise", "all"):
            run_reference_pairwise(prefix + "_inclusionCounts.tsv", prefix + "_allClusters.tsv",
                                   os.path.join(exp, f"pairwise_{mode}.tsv"), mode)
    relpath_manifest(manifest, d)
    with open(os.path.join(d, "quant_args.json"), "w") as f:
        json.dump(quant_over, f)


# --- every sample-file type on every filter boundary (SPLICEDICE.py:162-228, 257-295) -----------
MIXED_VARIANTS = {
    "default": {},
    "strict": dict(noMultimap=True, minOverhang=8, minEntropy=1.5, lowCoverageNan=True, minUnique=7),
    "lengths": dict(maxLength=1000, minLength=100, lowCoverageNan=True),
}


def write_mixed_inputs(dirname, seed=31):
    """A manifest of SJ.out.tab, tagged ``splicedicebed``, plain BED and leafcutter files (plus a
    ``.bam`` and an unknown suffix, which quant opens and ignores) drawn from a small coordinate
    pool so that the same junction recurs across files with scores either side of minUnique, with
    lengths on both sides of every length bound of MIXED_VARIANTS (strict for SJ, inclusive for
    BED), every SJ strand / motif code, overhang / entropy tags on their thresholds, annotated
    and unannotated tags, '.' strands and repeated lines (the last one wins)."""
    os.makedirs(dirname, exist_ok=True)
    rng = np.random.default_rng(seed)
    spans = [49, 50, 51, 99, 100, 101, 999, 1000, 1001, 49999, 50000, 50001, 300, 700, 4000]
    chroms = ["chr1", "chr10", "chr2"]
    starts = [1000, 1040, 1300, 2000, 2050, 60000]

    def draw():
        c = chroms[int(rng.integers(0, 3))]
        left = starts[int(rng.integers(0, len(starts)))]
        return c, left, left + spans[int(rng.integers(0, len(spans)))]

    def sj_lines(n):
        out = []
        for _ in range(n):
            c, left, right = draw()
            strand = int(rng.choice([0, 1, 1, 2, 2]))
            motif = int(rng.choice([0, 1, 1, 2, 2, 3, 4, 5, 6]))
            uniq, multi = int(rng.integers(0, 10)), int(rng.integers(0, 6))
            out.append(f"{c}\t{left + 1}\t{right}\t{strand}\t{motif}\t{int(rng.integers(0, 2))}\t{uniq}\t{multi}\t"
                       f"{int(rng.integers(1, 40))}\n")
        return out

    def tagged_lines(n):
        out = []
        for _ in range(n):
            c, left, right = draw()
            e1, e2 = (float(rng.choice([0.5, 0.99, 1.0, 1.01, 1.49, 1.5, 1.51, 2.2])) for _ in range(2))
            over = int(rng.choice([3, 4, 5, 6, 7, 8, 9, 20]))
            annot = str(rng.choice(["?", "?", "?", "GENE1", "ENSG0001"]))
            strand = str(rng.choice(["+", "+", "-", "-", "."]))
            out.append(f"{c}\t{left}\t{right}\te:{e1:0.02f}:{e2:0.02f};o:{over};m:GT_AG;a:{annot}\t"
                       f"{int(rng.integers(0, 14))}\t{strand}\n")
        return out

    def plain_lines(n, tag):
        out = []
        for k in range(n):
            c, left, right = draw()
            strand = str(rng.choice(["+", "+", "-", "-", "."]))
            out.append(f"{c}\t{left}\t{right}\t{tag}{k}\t{int(rng.integers(0, 14))}\t{strand}\n")
        return out

    files = {
        "a.SJ.out.tab": sj_lines(260), "b.SJ.out.tab": sj_lines(260),
        "c.bed": tagged_lines(260), "d.bed": tagged_lines(260),
        "e.junc.bed": plain_lines(220, "j"), "f.leafcutter.junc": plain_lines(220, "lc"),
        "g.bam": ["not a junction file\n"], "h.txt": plain_lines(20, "u"),
    }
    # repeated lines: the later score replaces the earlier one (SPLICEDICE.py:274)
    for name in ("a.SJ.out.tab", "c.bed", "e.junc.bed"):
        files[name] += files[name][:25][::-1]
    man = []
    for k, (name, lines) in enumerate(files.items()):
        path = os.path.join(dirname, name)
        with open(path, "w") as f:
            f.writelines(lines)
        man.append(f"s{k}_{name.split('.')[0]}\t{path}\tmeta{k}\tcond{k % 2}\n")
    mpath = os.path.join(dirname, "manifest.txt")
    with open(mpath, "w") as f:
        f.writelines(man)
    return mpath


def gen_mixed_case():
    d = case_dir("mixed_formats")
    manifest = write_mixed_inputs(os.path.join(d, "input"))
    for tag, over in MIXED_VARIANTS.items():
        exp = os.path.join(d, f"expected_{tag}")
        os.makedirs(exp)
        run_reference_quant(manifest, os.path.join(exp, "ref"), **over)
    relpath_manifest(manifest, d)
    with open(os.path.join(d, "variants.json"), "w") as f:
        json.dump(MIXED_VARIANTS, f, indent=1)


def gen_inmemory_quant(name, junctions, n_samples, seed, zero_frac=0.3, n_low=60):
    """getClusters + calculatePsi on injected data; adjacency stored as row lists."""
    rng = np.random.default_rng(seed)
    J = len(junctions)
    counts = synth.counts_host(seed, 0, J, n_samples)
    counts[rng.random((J, n_samples)) < zero_frac] = 0
    low = sorted({(int(a), int(b)) for a, b in zip(rng.integers(0, J, n_low),
                                                   rng.integers(0, n_samples, n_low))})
    clusters, psi = rh.ref_calculate_psi(junctions, counts.astype(np.float32), low)
    _, psi_nolow = rh.ref_calculate_psi(junctions, counts.astype(np.float32), None)
    index = {j: i for i, j in enumerate(sorted(clusters))}
    adj = [[index[o] for o in clusters[j]] for j in sorted(clusters)]
    row_ptr = np.zeros(J + 1, dtype=np.int64)
    row_ptr[1:] = np.cumsum([len(a) for a in adj])
    col_idx = np.array([c for a in adj for c in a], dtype=np.int64)
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        chrom=np.array([j[0] for j in junctions]), strand=np.array([j[3] for j in junctions]),
        start=np.array([j[1] for j in junctions], dtype=np.int64),
        end=np.array([j[2] for j in junctions], dtype=np.int64),
        counts=counts.astype(np.int32), low=np.array(low, dtype=np.int64).reshape(-1, 2),
        row_ptr=row_ptr, col_idx=col_idx,
        psi_bits=psi.view(np.uint32), psi_nolow_bits=psi_nolow.view(np.uint32))


def gen_fisher_tables():
    from scipy.stats import fisher_exact
    rng = np.random.default_rng(20261018)
    tabs = []
    for scale, n in ((3, 300), (10, 400), (50, 600), (200, 600), (1000, 500), (5000, 200)):
        tabs.append(rng.integers(0, scale, size=(n, 4)))
    nb = rng.negative_binomial(2, 0.02, size=(800, 2))            # config-3 like tables
    ex = rng.negative_binomial(9, 0.02, size=(800, 2))
    tabs.append(np.concatenate([nb, ex], axis=1))
    sym = []
    for _ in range(400):                                            # exact mirror ties
        a, b = (int(v) for v in rng.integers(0, 80, 2))
        sym.append([a, b, b, a])
        c = int(rng.integers(0, a + b + 1))
        sym.append([a, b, c, a + b - c])                            # n1 == n2
        sym.append([a, c, b, a + b - c if a + b - c >= 0 else 0])
    tabs.append(np.array(sym))
    tabs.append(np.array([[0, 0, 0, 0], [0, 0, 5, 7], [5, 0, 7, 0], [0, 5, 0, 7], [5, 7, 0, 0],
                          [1, 0, 0, 1], [0, 1, 1, 0], [1, 1, 1, 1], [2, 3, 3, 4], [3000, 1, 1, 3000],
                          [100, 0, 0, 100], [0, 100, 100, 0], [1, 2000, 2000, 1], [7, 7, 7, 7],
                          [10, 1000, 1000, 10], [20000, 19000, 18000, 21000]]))
    tabs = np.concatenate(tabs).astype(np.int64)
    p = np.array([fisher_exact([[t[0], t[1]], [t[2], t[3]]])[1] for t in tabs], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "fisher_tables.npz"), tables=tabs, p=p)
    # exact rational sums for a small-count subset (independent of Boost)
    import mpmath
    mpmath.mp.dps = 60
    sub = tabs[:700]
    exact = []
    for a, b, c, d in sub.tolist():
        n1, n2, n = a + b, c + d, a + c
        if n1 == 0 or n2 == 0 or n == 0 or b + d == 0:
            exact.append(1.0)
            continue
        lo, hi = max(0, n - n2), min(n1, n)
        w = [mpmath.binomial(n1, x) * mpmath.binomial(n2, n - x) for x in range(lo, hi + 1)]
        wa = w[a - lo]
        tot = mpmath.fsum(w)
        sel = mpmath.fsum(v for v in w if v <= wa)                 # exact-arithmetic two-sided rule
        exact.append(float(sel / tot))
    np.savez_compressed(os.path.join(GOLD, "fisher_exact_small.npz"), tables=sub,
                        p=np.array(exact))


def gen_ir_case():
    """ir_table on a tiny coverage directory (np.float shimmed; -r -j, threshold as float)."""
    d = case_dir("ir_small")
    rng = np.random.default_rng(5)
    junctions = sorted(synth.junction_tuples(60, 11))
    J, S = len(junctions), 4
    counts = synth.counts_host(3, 0, J, S)
    counts[rng.random((J, S)) < 0.25] = 0
    clusters = rh.ref_get_clusters(junctions)
    name = lambda j: f"{j[0]}:{j[1]}-{j[2]}:{j[3]}"  # noqa: E731
    samples = [f"ir{i}" for i in range(S)]
    with open(os.path.join(d, "counts.tsv"), "w") as f:
        f.write("cluster\t" + "\t".join(samples) + "\n")
        for i, j in enumerate(junctions):
            f.write(name(j) + "\t" + "\t".join(f"{x:.0f}" for x in counts[i]) + "\n")
    with open(os.path.join(d, "clusters.tsv"), "w") as f:
        for j in junctions:
            f.write(name(j) + "\t" + ",".join(name(o) for o in clusters[j]) + "\n")
    cov = os.path.join(d, "cov")
    os.makedirs(cov)
    med = rng.poisson(4, size=(J, S)).astype(float)
    med[rng.random((J, S)) < 0.1] = 0.0
    pts = rng.poisson(6, size=(J, S, 5)).astype(float) + 1.0
    for s, smp in enumerate(samples):
        with open(os.path.join(cov, f"{smp}_intron_coverage.txt"), "w") as f:
            for i, j in enumerate(junctions):
                pos = ",".join(str(j[1] + k) for k in range(5))
                c5 = ",".join(f"{v:g}" for v in pts[i, s])
                f.write(f"{j[0]}\t{j[1]}\t{j[2]}\t.\t{med[i, s]:g}\t{j[3]}\t{pos}\t{c5}\n")
    mod = rh.load("ir_table")
    args = rh._Args(allJunctions=True, makeRSDtable=True, singleJunctionCalculation=False,
                    RSDthreshold=1.0)
    cts = mod.getInclusionCounts(os.path.join(d, "counts.tsv"))
    cl = mod.getClusters(os.path.join(d, "clusters.tsv"))
    kept, IR, RSD = _quiet(mod.calculateIR, samples, cov, cts, cl, None, args)
    os.makedirs(os.path.join(d, "expected"))
    mod.writeIRtable(samples, os.path.join(d, "expected", "ref"), kept, IR)
    mod.writeRSDtable(samples, os.path.join(d, "expected", "ref"), kept, RSD)
    args.singleJunctionCalculation = True
    cts = mod.getInclusionCounts(os.path.join(d, "counts.tsv"))
    kept, IR, RSD = _quiet(mod.calculateIR, samples, cov, cts, None, None, args)
    mod.writeIRtable(samples, os.path.join(d, "expected", "ref_single"), kept, IR)
    with open(os.path.join(d, "samples.json"), "w") as f:
        json.dump(samples, f)


def main():
    if not rh.available():
        raise SystemExit("reference not found; golden fixtures can only be made in the build container")
    os.makedirs(GOLD, exist_ok=True)
    import scipy

    # 1. SURVEY §4 known-answer vector through the real file-based CLI path
    js = [(c, l, r, s) for c, l, r, s, _ in SURVEY_ROWS]
    counts = np.array([k for *_, k in SURVEY_ROWS], dtype=np.int64)
    gen_cli_case("survey_vector", js, counts, quant_over=dict(drim=True))

    # 2. config-1 analogue, small: 8 samples x ~400 junctions (+ adversarial structures)
    rng = np.random.default_rng(8)
    js = sorted(set(synth.junction_tuples(300, 4)) | set(synth.adversarial_tuples(2)))
    js = [j for j in js if 50 <= j[2] - j[1] <= 50000]
    counts = synth.counts_host(12, 0, len(js), 8).astype(np.int64)
    counts[rng.random(counts.shape) < 0.35] = 0
    gen_cli_case("cli_8x", js, counts, pairwise=False)
    gen_cli_case("cli_8x_lownan", js, counts, quant_over=dict(lowCoverageNan=True, minUnique=12),
                 pairwise=False)
    # a 5-sample slice for pairwise (10 pairs x ~80 events)
    js5 = js[:80]
    gen_cli_case("cli_5x_pairwise", js5, counts[:80, :5])

    # 2b. all four sample-file types x three flag sets
    gen_mixed_case()

    # 3. in-memory getClusters + calculatePsi
    gen_inmemory_quant("quant_adversarial", synth.adversarial_tuples(1), 6, seed=21)
    gen_inmemory_quant("quant_synth_3k", synth.junction_tuples(3000, 9), 16, seed=22)

    # 4. Fisher tables vs scipy.stats.fisher_exact
    gen_fisher_tables()

    # 5. ir_table
    gen_ir_case()

    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump({"generator": "oracle/gen_golden.py", "reference": "BrooksLabUCSC/splicedice @ /root/reference",
                   "numpy": np.__version__, "scipy": scipy.__version__,
                   "note": "BH correction in pairwise_{pairwise,all}.tsv comes from the harness stand-in "
                           "for the absent statsmodels (oracle_np.bh_adjust, pinned to "
                           "scipy.stats.false_discovery_control within 2 ulp: tests/test_bh_pin.py)"}, f, indent=1)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
