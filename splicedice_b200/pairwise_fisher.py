"""``splicedice pairwise`` on a B200: per-event Fisher exact test for every sample pair.

Host mirror of /root/reference/splicedice/pairwise_fisher.py (same flags, inputs and output
file).  The hot loop (pairwise_fisher.py:154-180) runs on the GPU: exclusion counts with
sd_quant_ps over the cluster CSR, then sd_fisher_pairwise for the [events x pairs] p-values.
Benjamini-Hochberg runs on the device too (sd_bh_adjust); the writer is the native host
formatter.  ``--chi2`` is not implemented (out of scope: BASELINE.json names the Fisher path
only) and is rejected loudly.
"""
from __future__ import annotations

import numpy as np

from . import junctions as jn


def getClusters(filename, filter_list=None):
    """event -> list of mutually exclusive events.  Whitespace split; a line that is not exactly
    two fields is an event with no partners (pairwise_fisher.py:26-43)."""
    clusters = {}
    with open(filename) as handle:
        for line in handle:
            fields = line.rstrip().split()
            if len(fields) == 2:
                clusters[fields[0]] = fields[1].split(",")
            else:
                clusters[line.strip()] = []
    return clusters


def getEventCounts(filename, filter_list=None):
    """(sample names, event names, float64 counts); with a filter only listed events are kept.
    Parsed natively (sd_host_table_read); same results as the reference's per-line reader."""
    from . import textio
    header, events, counts = textio.read_table(filename)
    samples = header.rstrip().split("\t")[1:]
    if filter_list is not None:
        keep = [i for i, e in enumerate(events) if e in filter_list]
        events = [events[i] for i in keep]
        counts = counts[keep]
    return samples, events, counts


def fdr_bh(p, device=0):
    """Benjamini-Hochberg adjusted p-values of a 1-D array, as statsmodels
    multipletests(method='fdr_bh')[1] (pairwise_fisher.py:185,190), computed on the GPU
    (sd_bh_adjust: radix sort, p / (rank / n), running minimum from the right, clip at 1)."""
    from . import ops
    import torch
    p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1)
    if p.size == 0:
        return p.copy()
    d = torch.from_numpy(p).to(torch.device("cuda", device)).reshape(-1, 1)
    return ops.bh_adjust(d, "all", out=d).reshape(-1).cpu().numpy()


def sample_pairs(n_samples):
    return [(i, j) for i in range(n_samples - 1) for j in range(i + 1, n_samples)]


def pairwise_pvalues(events, counts, clusters, device=0, correction="none", gpus=1):
    """float64[len(events), n_pairs] two-sided Fisher p-values, pair order as ``sample_pairs``;
    ``correction`` = 'none' | 'pairwise' | 'all' applies Benjamini-Hochberg on the device before
    the matrix comes back (pairwise_fisher.py:182-191)."""
    from . import ops
    import torch
    n_events, n_samples = counts.shape if counts.ndim == 2 else (0, 0)
    pairs = sample_pairs(n_samples)
    if n_events == 0 or not pairs:
        return np.zeros((n_events, len(pairs)))
    as_int = counts.astype(np.int64)
    if not np.array_equal(as_int, counts):
        # the reference sums the float rows first and lets scipy truncate the summed table
        # (pairwise_fisher.py:158-165), so fractions would add up differently: refuse rather than guess
        raise ValueError("inclusion counts must be integers (the table `quant` writes); fractional counts are "
                         "not supported on the B200 path")
    if (as_int < 0).any():
        raise ValueError("All values in `table` must be nonnegative.")
    if (as_int >= 2 ** 31).any():
        raise ValueError("counts of 2^31 and above are not supported")
    row_ptr, col_idx = jn.csr_from_named_lists(events, clusters, "isin")
    if gpus > 1:                                    # one worker process per GPU (row slabs; NCCL for the correction)
        from . import multigpu
        return multigpu.pairwise(as_int, row_ptr, col_idx, len(pairs), correction=correction, n_gpus=gpus)
    dev = torch.device("cuda", device)
    S = n_samples
    buf = torch.zeros((n_events, (S + 3) // 4 * 4), dtype=torch.int32, device=dev)
    buf[:, :S] = torch.from_numpy(as_int.astype(np.int32)).to(dev)
    inc = buf[:, :S]
    exc = ops.quant_ps(inc, row_ptr, col_idx, want_f32=False, want_exc=True)["exc"]
    pa = np.array([a for a, _ in pairs], dtype=np.int32)
    pb = np.array([b for _, b in pairs], dtype=np.int32)
    # the host filled the matrix, so it can vouch for a bound on inc + exc: no reduction / sync on the device
    bound = int(as_int.max(initial=0)) * (1 + int(np.diff(row_ptr).max(initial=0)))
    p = ops.fisher_pairwise(inc, exc, pa, pb, max_cell_bound=bound)
    if correction != "none":
        ops.bh_adjust(p, correction, out=p)
    return p.cpu().numpy()


def add_parser(parser):
    parser.add_argument("--inclusionSPLICEDICE", type=str, required=True,
                        help="inclusionCounts.tsv written by quant")
    parser.add_argument("-c", "--clusters", type=str, required=True, help="allClusters.tsv written by quant")
    parser.add_argument("--chi2", action="store_true", default=False,
                        help="(reference flag) chi-squared instead of Fisher -- not available on the GPU path")
    parser.add_argument("--multiple_test_correction", default="pairwise", choices=["pairwise", "all", "none"],
                        help="Benjamini-Hochberg per sample pair (default), over all p-values, or none")
    parser.add_argument("-f", "--filter_list", help="text file of events to analyse, one per line")
    parser.add_argument("-o", "--output", default="pairwise.tsv", help="output TSV")
    parser.add_argument("--device", type=int, default=0, help="CUDA device ordinal")
    parser.add_argument("--gpus", type=int, default=1,
                        help="GPUs: one worker process per GPU on devices 0..N-1, events cut into row slabs")


def run_with(args):
    if args.chi2:
        raise NotImplementedError("--chi2 is outside the B200 hot path (Fisher only); use the reference for it")
    filter_list = None
    if args.filter_list is not None:
        with open(args.filter_list) as handle:
            filter_list = {line.rstrip() for line in handle}
    samples, events, counts = getEventCounts(args.inclusionSPLICEDICE, filter_list)
    print("Counts loaded from", args.inclusionSPLICEDICE, "...")
    clusters = getClusters(args.clusters, filter_list)
    print("Clusters loaded from", args.clusters, "...")
    pairs = sample_pairs(len(samples))
    columns = [f"{samples[a]}_{samples[b]}" for a, b in pairs]
    print("Analyzing pairs:")
    print(",".join(columns))
    parray = pairwise_pvalues(events, counts, clusters, getattr(args, "device", 0),
                              correction=args.multiple_test_correction, gpus=getattr(args, "gpus", 1))
    print(f"[{len(events)} / {len(events)}] events analyzed...")
    from . import textio
    textio.write_matrix(args.output, "clusterID\t" + "\t".join(columns) + "\n", events,
                        np.ascontiguousarray(parray, dtype=np.float64), repr_floats=True)


if __name__ == "__main__":
    import argparse
    cli = argparse.ArgumentParser(description="pairwise Fisher exact tests (B200)")
    add_parser(cli)
    run_with(cli.parse_args())
