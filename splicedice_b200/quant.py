"""``splicedice quant`` on a B200: host side of the reference's SPLICEDICE.py.

Same command-line flags, manifest / sample-file formats and output files as
/root/reference/splicedice/SPLICEDICE.py; the two compute stages run on the GPU through the
C-ABI:

    getClusters   (SPLICEDICE.py:230-255, row index :96)  -> sd_cluster_build / sd_cluster_fill
    calculatePsi  (SPLICEDICE.py:297-310)                 -> sd_quant_ps_host

File reading and the TSV writers stay host Python (they are I/O).
Counts are held as int32 (the reference uses float32, exact below 2^24; scores of 2^24 and above
are kept exact here where the reference would round them).
"""
from __future__ import annotations

import os
from time import time

import numpy as np

from . import junctions as jn

BED_LIKE = ("bed", "splicedicebed", "leafcutter")
_STRAND_OF_SJ = {"0": "0", "1": "+", "2": "-", "+": "+", "-": "-"}
_MOTIF_SETS = {"gtag_only": {1, 2}, "gc_at": {1, 2, 3, 4, 5}, "all": {0, 1, 2, 3, 4, 5, 6}}


class Sample:
    """One manifest row; the file type is sniffed from the file name (SPLICEDICE.py:22-36)."""

    def __init__(self, manifest_row):
        self.name = manifest_row[0]
        self.filename = manifest_row[1]
        upper = self.filename.upper()
        if upper.endswith(".BED"):
            self.type = "bed"
            with open(self.filename) as handle:
                tags = handle.readline().split("\t")[3].split(";")
            if tags[0].startswith("e:") and tags[1].startswith("o:"):
                self.type = "splicedicebed"
        elif upper.endswith("SJ.OUT.TAB"):
            self.type = "SJ"
        elif upper.endswith(".BAM"):
            self.type = "bam"
        elif upper.endswith("LEAFCUTTER.JUNC"):
            self.type = "leafcutter"
        else:
            self.type = "unknown"
        self.metadata = manifest_row[2]
        self.condition = manifest_row[3]


class _NativeIngest:
    """sd_ingest_* handle: the junction union (pass 1) and the count matrix (pass 2)."""

    _TYPE_CODE = {"SJ": 0, "splicedicebed": 1, "bed": 2, "leafcutter": 2}

    def __init__(self, manifest, args):
        import ctypes
        from . import native
        self._ct, self._native = ctypes, native
        self.lib = native.load()
        self.handle = ctypes.c_void_p(self.lib.sd_ingest_create())
        self.manifest = manifest
        motifs = _MOTIF_SETS[args.filter]
        self.filter = native.QuantFilter(args.maxLength, args.minLength, args.minOverhang, args.minUnique,
                                         int(bool(args.noMultimap)), int(bool(args.lowCoverageNan)),
                                         sum(1 << m for m in motifs), 0, float(args.minEntropy))
        n = len(manifest)
        self.paths = (ctypes.c_char_p * n)(*[os.fsencode(s.filename) for s in manifest])
        self.types = np.array([self._TYPE_CODE.get(s.type, -1) for s in manifest], dtype=np.int32)
        self.threads = int(getattr(args, "threads", 0) or 0)

    def __del__(self):
        try:
            self.lib.sd_ingest_destroy(self.handle)
        except Exception:
            pass

    def _check(self, name, rc):
        if rc != 0:
            raise ValueError(f"{name}: {self._native.last_error()}")

    def collect(self):
        ct, native = self._ct, self._native
        self._check("sd_ingest_collect", self.lib.sd_ingest_collect(
            self.handle, len(self.manifest), self.paths, native.ptr(self.types), ct.byref(self.filter), self.threads))
        n = self.lib.sd_ingest_junction_count(self.handle)
        chrom = np.empty(n, dtype=np.int32); left = np.empty(n, dtype=np.int32)
        right = np.empty(n, dtype=np.int32); strand = np.empty(n, dtype=np.int8)
        self._check("sd_ingest_export", self.lib.sd_ingest_export(
            self.handle, native.ptr(chrom), native.ptr(left), native.ptr(right), native.ptr(strand)))
        names = [self.lib.sd_ingest_chrom_name(self.handle, i).decode() for i in range(self.lib.sd_ingest_chrom_count(self.handle))]
        return {(names[c], l, r, chr(s)) for c, l, r, s in zip(chrom.tolist(), left.tolist(), right.tolist(), strand.tolist())}

    def counts(self, rows):
        """rows: junction tuples in output-row order -> (int32[J, S], low mask or None)."""
        ct, native = self._ct, self._native
        n = len(rows)
        chrom_names = sorted({j[0] for j in rows})
        lut = {c: i for i, c in enumerate(chrom_names)}
        cnames = (ct.c_char_p * max(len(chrom_names), 1))(*[c.encode() for c in chrom_names])
        chrom_of = np.fromiter((lut[j[0]] for j in rows), dtype=np.int32, count=n)
        left = np.fromiter((j[1] for j in rows), dtype=np.int32, count=n)
        right = np.fromiter((j[2] for j in rows), dtype=np.int32, count=n)
        strand = np.fromiter((ord(j[3]) for j in rows), dtype=np.int8, count=n)
        row = np.arange(n, dtype=np.int32)
        self._check("sd_ingest_index", self.lib.sd_ingest_index(
            self.handle, n, cnames, native.ptr(chrom_of), native.ptr(left), native.ptr(right), native.ptr(strand),
            native.ptr(row)))
        S = len(self.manifest)
        counts = np.zeros((n, S), dtype=np.int32)
        mask = np.zeros((n, S), dtype=np.uint8) if self.filter.low_coverage_nan else None
        samples = np.arange(S, dtype=np.int32)
        self._check("sd_ingest_counts", self.lib.sd_ingest_counts(
            self.handle, S, self.paths, native.ptr(self.types), native.ptr(samples), ct.byref(self.filter),
            native.ptr(counts), S, native.ptr(mask), S, self.threads))
        return counts, mask


class LowCells:
    """The reference's ``low`` list of (row, sample) cells (SPLICEDICE.py:275-276) held as the
    native reader's uint8 mask; the tuples are only materialised if somebody iterates."""

    def __init__(self, mask):
        self.mask = mask

    def __len__(self):
        return int(np.count_nonzero(self.mask))

    def __bool__(self):
        return bool(self.mask.any())

    def __iter__(self):
        return (tuple(x) for x in np.argwhere(self.mask).tolist())


class Timer:
    """Stage timer printing the reference's ``[h:mm:ss.ss]`` stamps (SPLICEDICE.py:48-66)."""

    def __init__(self):
        self.start = self.checkpoint = time()

    @staticmethod
    def _fmt(seconds):
        return f"[{int(seconds // 3600)}:{int((seconds % 3600) // 60):02d}:{seconds % 60:02.2f}]"

    def total(self):
        return self._fmt(time() - self.start)

    def check(self):
        now = time()
        out = self._fmt(now - self.checkpoint)
        self.checkpoint = now
        return out


class SPLICEDICE:
    """The quant pipeline.  Constructing it runs every stage, like the reference class; pass
    ``run=False`` to drive the stages by hand (tests, library use)."""

    def __init__(self, manifestFilename, outputPrefix, args, device=0, run=True, native_io=True, gpus=1):
        self.args = args
        self.gpus = gpus
        self.native_io = native_io
        self.manifestFilename = manifestFilename
        self.outputPrefix = outputPrefix
        self.device = device
        self._csr = None
        self._clusters = None
        if run:
            self.run()

    # ---------------------------------------------------------------- pipeline ----------
    def run(self):
        timer = Timer()
        print("Parsing manifest...")
        self.manifest = self.parseManifest()
        print("\tDone", timer.check())
        print(f"Getting all junctions from {len(self.manifest)} files...")
        self.junctions = self.getAllJunctions()
        print("\tDone", timer.check())
        print(f"Finding clusters from {len(self.junctions)} junctions...")
        self.getClusters()
        print("\tDone", timer.check())
        print("Writing cluster file...")
        self.writeClusters()
        print("\tDone", timer.check())
        print("Writing junction bed file...")
        self.writeJunctionBed()
        print("\tDone", timer.check())
        print("Gathering junction counts...")
        self.counts, self.low = self.getJunctionCounts()
        print("\tDone", timer.check())
        print("Writing inclusion counts...")
        self.writeInclusions()
        print("\tDone", timer.check())
        print("Calculating PS values...")
        self.psi = self.calculatePsi()
        print("\tDone", timer.check())
        print("Writing PS values...")
        self.writeAllpsi()
        print("\tDone", timer.check())
        if getattr(self.args, "npz", False):
            print("Writing PS matrix (npz)...")
            self.writeNpz()
            print("\tDone", timer.check())
        if self.args.drim:
            print("Writing drim table...")
            self.writeDrimTable()
            print("\tDone", timer.check())
        print("All done", timer.total())

    def parseManifest(self):
        """name, path, metadata, condition per line; no validation (SPLICEDICE.py:134-144)."""
        with open(self.manifestFilename) as handle:
            return [Sample(line.rstrip().split("\t")) for line in handle]

    # ---------------------------------------------------------------- junction union ----
    def _admit_sj(self, row):
        a = self.args
        left, right = int(row[1]) - 1, int(row[2])
        strand = _STRAND_OF_SJ[row[3]]
        score = int(row[6]) if a.noMultimap else int(row[6]) + int(row[7])
        span = right - left
        if (a.minLength < span < a.maxLength and strand != "0" and score >= a.minUnique
                and int(row[4]) in self._motifs):
            return (row[0], left, right, strand)
        return None

    def _admit_tagged_bed(self, row):
        """bam_to_junc_bed output: the filters apply only to unannotated junctions (a:?)."""
        a = self.args
        score = int(row[4])
        tags = [t.split(":") for t in row[3].split(";")]
        left, right = int(row[1]), int(row[2])
        if tags[3][1] == "?":
            span = right - left
            if (score < a.minUnique or span > a.maxLength or span < a.minLength
                    or int(tags[1][1]) < a.minOverhang
                    or float(tags[0][1]) < a.minEntropy or float(tags[0][2]) < a.minEntropy):
                return None
        return (row[0], left, right, row[5]) if row[5] in ("+", "-") else None

    def _admit_plain_bed(self, row):
        a = self.args
        if int(row[4]) < a.minUnique:
            return None
        left, right = int(row[1]), int(row[2])
        span = right - left
        if span > a.maxLength or span < a.minLength:
            return None
        return (row[0], left, right, row[5]) if row[5] in ("+", "-") else None

    def getAllJunctions(self):
        """Union over samples of the junctions that pass that sample's filter
        (SPLICEDICE.py:147-228).  ``.bam`` / unknown-suffix samples contribute nothing.
        Native multi-threaded reader (sd_ingest_collect) unless ``native_io`` is off."""
        if not self.native_io:
            return self._getAllJunctions_py()
        ing = self._ingest = _NativeIngest(self.manifest, self.args)
        return ing.collect()

    def _getAllJunctions_py(self):
        """The same union with plain python line parsing (kept as the cross-check of the native reader)."""
        self._motifs = _MOTIF_SETS[self.args.filter]
        admit_by_type = {"SJ": self._admit_sj, "splicedicebed": self._admit_tagged_bed,
                         "bed": self._admit_plain_bed, "leafcutter": self._admit_plain_bed}
        found = set()
        for sample in self.manifest:
            with open(sample.filename) as handle:      # opened even when ignored, as the reference does
                admit = admit_by_type.get(sample.type)
                if admit is None:
                    continue
                for line in handle:
                    j = admit(line.rstrip().split("\t"))
                    if j is not None:
                        found.add(j)
        return found

    # ---------------------------------------------------------------- clusters (GPU) ----
    def getClusters(self):
        """Overlap adjacency on the device.  Fills ``junctionIndex`` (tuple -> output row, the
        order of ``sorted(junctions)``) and the CSR; returns the reference-shaped dict lazily via
        ``self.clusters``."""
        from . import ops
        table = jn.JunctionTable(self.junctions)
        built = ops.cluster_build(*table.arrays(), device=self.device)
        out_row = built["out_row"].cpu().numpy()
        self._rows = jn.rows_in_output_order(table.tuples, out_row)
        self._csr = (built["row_ptr"].cpu().numpy(), built["col_idx"].cpu().numpy())
        self.n_components = built["n_comp"]
        self.junctionIndex = {j: r for r, j in enumerate(self._rows)}
        self._clusters = None
        return self.clusters

    @property
    def clusters(self):
        if self._clusters is None and self._csr is not None:
            self._clusters = jn.adjacency_dict(self._rows, *self._csr)
        return self._clusters

    # ---------------------------------------------------------------- counts -------------
    def getJunctionCounts(self):
        """int32[J, S] of scores (last duplicate line wins) and the list of low cells
        (SPLICEDICE.py:257-295).  Native reader (sd_ingest_counts) unless ``native_io`` is off."""
        if not self.native_io:
            return self._getJunctionCounts_py()
        ing = getattr(self, "_ingest", None) or _NativeIngest(self.manifest, self.args)
        counts, mask = ing.counts(self._rows)
        return counts, (LowCells(mask) if mask is not None else [])

    def _getJunctionCounts_py(self):
        index = self.junctionIndex
        a = self.args
        counts = np.zeros((len(index), len(self.manifest)), dtype=np.int32)
        low = []
        for s, sample in enumerate(self.manifest):
            with open(sample.filename) as handle:
                if sample.type in BED_LIKE:
                    for line in handle:
                        row = line.rstrip().split("\t")
                        r = index.get((row[0], int(row[1]), int(row[2]), row[5]))
                        if r is None:
                            continue
                        score = int(row[4])
                        counts[r, s] = score
                        if a.lowCoverageNan and score < a.minUnique:
                            low.append((r, s))
                elif sample.type == "SJ":
                    for line in handle:
                        row = line.rstrip().split("\t")
                        r = index.get((row[0], int(row[1]) - 1, int(row[2]), _STRAND_OF_SJ[row[3]]))
                        if r is None:
                            continue
                        counts[r, s] = int(row[6]) if a.noMultimap else int(row[6]) + int(row[7])
        return counts, low

    # ---------------------------------------------------------------- PS (GPU) -----------
    def calculatePsi(self):
        """float32[J, S] PS in output-row order; NaN on zero coverage and, with
        --lowCoverageNan, on the low cells (SPLICEDICE.py:297-310)."""
        from . import ops
        mask = None
        if self.args.lowCoverageNan and isinstance(self.low, LowCells):
            mask = self.low.mask if self.low else None
        elif self.args.lowCoverageNan and self.low:
            mask = np.zeros(self.counts.shape, dtype=np.uint8)
            rows, cols = zip(*self.low)
            mask[list(rows), list(cols)] = 1
        row_ptr, col_idx = self._csr
        if self.counts.size == 0:
            return np.zeros(self.counts.shape, dtype=np.float32)
        if self.gpus > 1:                           # one worker process per GPU, row slabs at cluster boundaries
            from . import multigpu
            return multigpu.quant_ps(np.ascontiguousarray(self.counts, dtype=np.int32), row_ptr, col_idx,
                                     low_mask=mask, n_gpus=self.gpus)
        return ops.quant_ps_host(np.ascontiguousarray(self.counts, dtype=np.int32), row_ptr, col_idx,
                                 low_mask=mask, device=self.device).numpy()

    # ---------------------------------------------------------------- writers ------------
    def junctionString(self, junction):
        return jn.junction_name(junction)

    def writeJunctionBed(self):
        with open(f"{self.outputPrefix}_junctions.bed", "w") as out:
            for chrom, left, right, strand in self._rows:
                out.write(f"{chrom}\t{left}\t{right}\t{chrom}:{left}-{right}:{strand}\t0\t{strand}\n")

    def writeClusters(self):
        names = [jn.junction_name(j) for j in self._rows]
        rp, ci = (x.tolist() for x in self._csr)
        with open(f"{self.outputPrefix}_allClusters.tsv", "w") as out:
            for r, name in enumerate(names):
                out.write(name + "\t" + ",".join(names[c] for c in ci[rp[r]:rp[r + 1]]) + "\n")

    def _write_matrix(self, path, matrix):
        """cluster<TAB>samples header, then name<TAB>values rows through the native formatter
        (byte-identical to the reference's f'{x:.0f}' / f'{x:.3f}' per cell)."""
        from . import textio
        names = [jn.junction_name(j) for j in self._rows]
        textio.write_matrix(path, "cluster\t" + "\t".join(s.name for s in self.manifest) + "\n", names, matrix)

    def writeInclusions(self):
        self._write_matrix(f"{self.outputPrefix}_inclusionCounts.tsv", np.ascontiguousarray(self.counts, dtype=np.int32))

    def writeAllpsi(self):
        self._write_matrix(f"{self.outputPrefix}_allPS.tsv", np.ascontiguousarray(self.psi, dtype=np.float32))

    def writeNpz(self):
        """``{prefix}_allPS.npz`` with ``cols`` (sample names), ``rows`` (junction names) and
        ``data`` (float32 PS, rows x cols) -- the layout findOutliers.py:117-118 loads and the
        reference's README still names for the downstream commands; nothing in the reference
        writes it, so this output is an addition behind ``--npz``."""
        np.savez(f"{self.outputPrefix}_allPS.npz",
                 cols=np.array([s.name for s in self.manifest]),
                 rows=np.array([jn.junction_name(j) for j in self._rows]),
                 data=np.ascontiguousarray(self.psi, dtype=np.float32))

    def writeDrimLine(self, i, junction, other, file):
        # the reference prints its float32 counts with astype(str): "34.0"
        values = "\t".join(f"{float(x)}" for x in self.counts[self.junctionIndex[other], :].tolist())
        print(f"cl_{i}_{self.junctionString(junction)}", f"{self.junctionString(other)}_{i}", values,
              sep="\t", file=file)

    def writeDrimTable(self):
        rp, ci = (x.tolist() for x in self._csr)
        with open(f"{self.outputPrefix}_drimTable.tsv", "w") as out:
            out.write("gene\tfeature_id\t" + "\t".join(s.name for s in self.manifest) + "\n")
            for i, junction in enumerate(self._rows):
                self.writeDrimLine(i, junction, junction, out)
                for c in ci[rp[i]:rp[i + 1]]:
                    self.writeDrimLine(i, junction, self._rows[c], out)


def add_parser(parser):
    parser.add_argument("--manifest", "-m", required=True,
                        help="tab-separated sample list: name, path, metadata, condition")
    parser.add_argument("--output_prefix", "-o", required=True, help="prefix of the output files")
    parser.add_argument("--maxLength", type=int, default=50000, help="longest junction kept")
    parser.add_argument("--minLength", type=int, default=50, help="shortest junction kept")
    parser.add_argument("--minOverhang", type=int, default=5, help="least read overhang supporting a junction")
    parser.add_argument("--drim", action="store_true", help="also write the DRIMSeq table")
    parser.add_argument("--noMultimap", action="store_true", help="count uniquely mapped reads only (SJ.out.tab)")
    parser.add_argument("--filter", default="gtag_only", choices=["gtag_only"], help="intron motifs kept")
    parser.add_argument("--minUnique", type=int, default=5, help="least score for a sample to admit a junction")
    parser.add_argument("--lowCoverageNan", action="store_true", help="NaN for cells scored below minUnique")
    parser.add_argument("--minEntropy", type=float, default=1, help="least Shannon diversity of read offsets")
    parser.add_argument("--device", type=int, default=0, help="CUDA device ordinal")
    parser.add_argument("--gpus", type=int, default=1,
                        help="GPUs for the PS stage: one worker process per GPU on devices 0..N-1, rows cut at cluster boundaries")
    parser.add_argument("--threads", type=int, default=0, help="host threads for file parsing (0 = all)")
    parser.add_argument("--pythonIO", action="store_true", help="read the sample files with the plain python parser")
    parser.add_argument("--npz", action="store_true",
                        help="also write {prefix}_allPS.npz (cols / rows / data, the matrix layout findOutliers reads)")


def run_with(args):
    SPLICEDICE(args.manifest, args.output_prefix, args, device=getattr(args, "device", 0),
               native_io=not getattr(args, "pythonIO", False), gpus=getattr(args, "gpus", 1))


if __name__ == "__main__":
    import argparse
    cli = argparse.ArgumentParser(description="SpliceDICE quant on a B200")
    add_parser(cli)
    run_with(cli.parse_args())
