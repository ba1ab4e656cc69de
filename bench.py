#!/usr/bin/env python
"""Benchmark of the SpliceDICE quant / pairwise hot path on B200 (one JSON line on stdout).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1

Headline workload (BASELINE.json configs[1]): quant PS, 1,000 samples x 400,000 junctions per GPU
(weak scaling: N GPUs hold N contiguous row slabs of an N x 400k-junction problem, cut where no
adjacency edge crosses -- splicedice_b200/sharding.py).  A step is one pass of the fused
exclusion-aggregation + PS kernel over the rank's slab.

  value        PS cells/s, counts and PS resident in HBM (CUDA events, max over ranks)
  e2e          the same metric through sd_quant_ps_host: pinned host counts in, pinned host PS
               out, H2D and D2H inside the timed region; e2e.fabric = the box's pinned-copy
               ceiling measured in the same run with no kernel (H2D alone, D2H alone, both at
               once, all ranks together), e2e.frac_of_fabric = duplex time / pipeline time
  roofline     8 algorithmic bytes per cell (4 B int32 count read + 4 B float32 PS written)
               over the kernel's average launch time, against MEASURED_PEAKS.json hbm_gbs;
               roofline.kernel is what the library says it launched (sd_quant_last_launch)
  cpu_baseline the UNMODIFIED reference (oracle/_ref, installed by oracle/build_ref.py) on the
               FULL 400,000 x 1,000 workload: SPLICEDICE.getClusters + row index + calculatePsi,
               fanned over all host cores in closed row slabs (the same measurement
               `--impl reference` makes), the single-core time of the same code beside it, and
               a bit-for-bit comparison of the reference's PS matrix with the GPU's
  cluster_build warm (median of 5) and cold (first call) time of the device cluster build
  variants     float64 PS (counts_to_ps) and the intron-retention ratio (configs[4])
  fisher       pairwise Fisher (configs[2]: 64 samples, 2,016 pairs x 200,000 junctions per GPU):
               tests/s, FP64 model, Benjamini-Hochberg, host-buffer e2e, reference CPU baseline
  strong       (N > 1) the FIXED configs cut over the N GPUs: configs[1] as one 400k x 1,000
               problem, configs[2] as one 200k x 64 problem -- per-rank time, max over ranks,
               next to the same problem on one GPU in the same run
  collectives  (N > 1) the product's NCCL paths on those slabs: all-gather of the PS row slabs
               (distributed.all_gather_rows) and the pairwise all-to-all -> sd_bh_adjust ->
               all-to-all (distributed.bh_columns_sharded), each with ms, GB/s and a bit-identity
               check against the single-GPU result
  tcga         configs[3]: 10,000 samples x 1,000,000 junctions cut over the N GPUs

`--impl reference` times only the unmodified reference on the host cores (no GPU).
Inputs are larger than L2 (1.6 GB in + 1.6 GB out per pass against 126 MB), so no flush is
needed between timed iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "quant_ps_cells_per_s"
UNIT = "cells/s"
SEED = 20261018


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--junctions", type=int, default=400_000, help="junctions per GPU")
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--fisher-junctions", type=int, default=200_000)
    ap.add_argument("--fisher-samples", type=int, default=64)
    ap.add_argument("--no-fisher", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the float64 PS / intron-retention timings")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling / collective sections (N > 1)")
    ap.add_argument("--no-tcga", action="store_true", help="skip configs[3] (10,000 x 1M)")
    ap.add_argument("--tcga-junctions", type=int, default=1_000_000)
    ap.add_argument("--tcga-samples", type=int, default=10_000)
    ap.add_argument("--cpu-fisher-events", type=int, default=6)
    ap.add_argument("--cpu-workers", type=int, default=0, help="processes of the reference fan-out (0 = all host cores)")
    ap.add_argument("--flags", type=int, default=0, help="sd_quant_ps flags (kernel variant / tile shape)")
    return ap.parse_args()


def base_config(args):
    """The part of `config` both arms print identically (the reference arm runs the same workload)."""
    return {"workload": f"quant PS: {args.samples} samples x {args.junctions} junctions per GPU "
                        f"({config_label(args.samples, args.junctions)})",
            "junctions_per_gpu": args.junctions, "samples": args.samples, "seed": SEED,
            "generator": "splicedice_b200/synth.py: junction_arrays(seed) + counter-based geometric counts(seed + 1)"}


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def committed_capture(name):
    """An ncu figure from a committed capture under profiles/ (a profiler cannot run inside the
    timed bench): returned with its source and the round it was taken in, so a reader can tell a
    capture of this round's kernel from an older one."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


def fisher_work(a, b, c, d, cut_bits=48):
    """(mean tail terms summed per test, mean hypergeometric support, trivial fraction) for tables
    [[a, b], [c, d]] -- the kernel's own stopping rule (sd_fisher_math.cuh: TailState), in numpy."""
    n1, n2, n = a + b, c + d, a + c
    N = n1 + n2
    trivial = (n1 == 0) | (n2 == 0) | (n == 0) | (b + d == 0)
    support = np.where(trivial, 0, np.minimum(n1, n) - np.maximum(0, n - n2) + 1)
    mode = np.floor((n + 1) * (n1 + 1) / (N + 2))
    known = trivial | (a == mode)
    swap = a > mode
    a2, b2 = np.where(swap, b, a), np.where(swap, a, b)
    c2, d2 = np.where(swap, d, c), np.where(swap, c, d)
    nn = np.where(swap, N - n, n)
    mode2 = np.where(swap, n1 - mode, mode)

    def bursts(p, q, u, v):
        P = np.ones_like(p); Q = np.ones_like(p); A = np.ones_like(p)
        num, s_, den, w = p * q, p + q - 1, (u + 1) * (v + 1), u + v + 3
        done = np.zeros(p.shape, bool); count = np.zeros(p.shape)
        for _ in range(2000):
            for _k in range(4):
                P = P * num; Q = Q * den; A = A * den + P
                num = num - s_; s_ = s_ - 2; den = den + w; w = w + 2
            big = Q > 2.0 ** 500
            P = np.where(big, P * 2.0 ** -500, P); Q = np.where(big, Q * 2.0 ** -500, Q); A = np.where(big, A * 2.0 ** -500, A)
            count = np.where(done, count, count + 1)
            # TailState::done(): integer comparison of the exponent words (top 32 bits)
            done |= (P.view(np.int64) >> 32) < (A.view(np.int64) >> 32) - (cut_bits << 20)
            if done.all():
                break
        return count
    with np.errstate(all="ignore"):
        near = bursts(a2, d2, b2, c2)
        g = np.minimum(2 * mode2 - a2, np.minimum(n1, nn))          # mirror point ~ first admitted far-side point
        far = bursts(n1 - g, nn - g, g, n2 - nn + g)
    terms = np.where(known, 0, 4 * (near + far))
    return float(terms.mean()), float(support.mean()), float(trivial.mean())


def config_label(samples, junctions):
    if samples == 1000 and junctions == 400_000:
        return "BASELINE.json configs[1], GTEx-scale"
    if samples == 10_000:
        return "BASELINE.json configs[3] shape, TCGA-scale"
    return "custom shape"


# ------------------------------------------------------------------------------------------
# host <-> device link, no kernel in the way
# ------------------------------------------------------------------------------------------
def fabric_probe(dev, h2d_bytes, d2h_bytes, chunk_mb=32, world=1, reps=3, buffers=None):
    """The host <-> device link with no kernel in the way: pinned-host -> device copies of
    ``h2d_bytes`` alone, device -> pinned-host copies of ``d2h_bytes`` alone, and both at once on
    two streams (what a perfectly overlapped host-buffer pipeline would take), in ``chunk_mb``
    pieces.  All ranks start together (barrier) and the slowest rank's time counts, so at N GPUs
    the figures are the aggregate the box sustains.  GB/s = bytes of ALL ranks / time.
    ``buffers`` = (pinned in, pinned out, device in, device out) uint8 tensors to reuse."""
    import torch
    import torch.distributed as dist
    chunk = chunk_mb << 20
    n_in, n_out = int(h2d_bytes), int(d2h_bytes)
    if buffers is None:
        h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
        d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
        d_out = torch.zeros(n_out, dtype=torch.uint8, device=dev)
    else:
        h_in, h_out, d_in, d_out = buffers
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(do_in, do_out):
        best = None
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if do_in:
                with torch.cuda.stream(s_in):
                    for o in range(0, n_in, chunk):
                        d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    for o in range(0, n_out, chunk):
                        h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
            s_in.synchronize(); s_out.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            best = dt if best is None else min(best, dt)
        return best
    t_in, t_out, t_both = run(True, False), run(False, True), run(True, True)
    return {"h2d_gbs": world * n_in / t_in / 1e9, "d2h_gbs": world * n_out / t_out / 1e9,
            "duplex_gbs": world * (n_in + n_out) / t_both / 1e9, "h2d_ms": t_in * 1e3, "d2h_ms": t_out * 1e3,
            "duplex_ms": t_both * 1e3, "bytes_per_gpu": [n_in, n_out], "chunk_mb": chunk_mb, "n_gpus": world,
            "what": "pinned cudaMemcpyAsync in chunks, no kernel, all ranks at once, best of %d; GB/s summed over ranks" % reps}


# ------------------------------------------------------------------------------------------
# CPU legs: the unmodified reference (oracle/_ref), else the loop-for-loop port
# ------------------------------------------------------------------------------------------
def host_workload(n_rows, n_samples):
    """sorted junction tuples + float32 counts of the headline workload, made on the host cores
    (the same matrix the device generator makes: tests/test_gpu_quant.py)."""
    from concurrent.futures import ThreadPoolExecutor
    from splicedice_b200 import synth
    rows = sorted(synth.junction_tuples(n_rows, SEED))
    counts = np.empty((n_rows, n_samples), dtype=np.float32)
    step = 10_000

    def fill(r):
        counts[r:r + step] = synth.counts_host(SEED + 1, r, min(step, n_rows - r), n_samples)
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        list(ex.map(fill, range(0, n_rows, step)))
    return rows, counts


class CpuQuant:
    """One getClusters + row index + calculatePsi pass of the reference per step(): the unmodified
    reference through oracle/ref_harness.RefQuantPool (workers forked once, outside the timed
    steps), or -- when no reference tree is installed -- the loop-for-loop port, one process."""

    def __init__(self, rows, counts_f32, workers, keep=None):
        from oracle import ref_harness
        self.rows, self.counts, self.keep = rows, counts_f32, keep
        self.kind = "reference" if ref_harness.available() else "port"
        self.pool = None
        if self.kind == "reference":
            self._quiet = ref_harness.warnings_off()
            self._quiet.__enter__()
            self.pool = ref_harness.RefQuantPool(rows, counts_f32, workers, keep)
            self.workers = self.pool.workers
        else:
            self.workers = 1

    def step(self):
        if self.pool is not None:
            return self.pool.step()[0]
        from oracle import ref_port
        t0 = time.perf_counter()
        adjacency = ref_port.sweep_clusters(self.rows)
        index = ref_port.row_index(adjacency)
        out = ref_port.psi_loop(adjacency, index, self.counts)
        if self.keep is not None:
            self.keep[:] = out
        return time.perf_counter() - t0

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self._quiet.__exit__(None, None, None)
            self.pool = None


def cpu_fisher_tests_per_s(n_events, n_samples):
    from oracle import ref_harness
    if ref_harness.available():
        with tempfile.TemporaryDirectory(prefix="sd_ref_pairwise_") as tmp, ref_harness.warnings_off():
            return ref_harness.ref_pairwise_tests_per_second(n_events, n_samples, SEED, tmp), "reference"
    from oracle import ref_port
    rng = np.random.default_rng(SEED)
    inc = rng.negative_binomial(2, 0.02, size=(n_events, n_samples)).astype(np.float64)
    names = np.array([f"chr1:{100 * i}-{100 * i + 150}:+" for i in range(n_events)])
    clusters = {names[i]: [names[j] for j in (i - 1, i + 1) if 0 <= j < n_events] for i in range(n_events)}
    t0 = time.perf_counter()
    ref_port.pairwise_loop(names, inc, clusters)
    return n_events * (n_samples * (n_samples - 1) // 2) / (time.perf_counter() - t0), "port"


def reference_sample_text(args, workers, kind):
    what = ("the UNMODIFIED reference (oracle/_ref: SPLICEDICE.getClusters + row index + calculatePsi, "
            "SPLICEDICE.py:230-255,96,297-310)" if kind == "reference" else
            "oracle/ref_port.py (loop-for-loop port; oracle/_ref is not installed)")
    fan = (f"fanned over {workers} processes in closed row slabs (oracle/ref_harness.closed_row_slabs)" if workers > 1
           else "one process")
    return f"the FULL {args.junctions} junctions x {args.samples} samples of one GPU's workload per step: {what}, {fan}"


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = args.cpu_workers or cores
    rows, counts = host_workload(args.junctions, args.samples)
    cells = args.junctions * args.samples
    cpu = CpuQuant(rows, counts, workers)
    kind, workers = cpu.kind, cpu.workers
    times = [cpu.step() for _ in range(args.warmup + args.steps)][args.warmup:]
    cpu.close()
    ms = 1e3 * float(np.mean(times))
    value = cells / (ms * 1e-3)
    sample = reference_sample_text(args, workers, kind)
    if args.gpus > 1:
        sample += f"; at {args.gpus} GPUs the workload is {args.gpus} such slabs and this is a 1/{args.gpus} sample of it"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 counts, f64 sums and divide, f32 PS (numpy)", "data": "synthetic",
        "config": base_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample,
                         "host_cores": cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    # stdout carries exactly one JSON line: anything libraries print there meanwhile (NCCL announces
    # its version on stdout at the first collective) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from splicedice_b200 import native, ops, sharding, synth

    ops.require_cuda()
    # host buffers of this rank on the GPU's own NUMA node (matters for the e2e figures at N > 1)
    from splicedice_b200 import distributed as sd_dist
    numa_cpus = sd_dist.bind_to_gpu_numa(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MAX)

    def sum_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.SUM)

    def all_true(flag):
        return bool(reduce_ranks(1.0 if flag else 0.0, dist.ReduceOp.MIN) > 0.5)

    def timed_graph(step, k, what):
        """ms per launch (this rank) of `k` back-to-back calls of `step`, captured into one CUDA graph
        so a busy host cannot open gaps between the launches; plain launches if capture fails."""
        graph = None
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for _ in range(k):
                        step()
            torch.cuda.current_stream().wait_stream(side)
            graph = g
        except Exception as exc_capture:            # noqa: BLE001 - any capture failure falls back to plain launches
            print(f"[bench] CUDA-graph capture unavailable for {what} ({exc_capture}); launching step by step", file=sys.stderr)
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        if graph is not None:
            graph.replay()
        else:
            for _ in range(k):
                step()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / k, graph is not None

    def timed_calls(fn, n):
        """median over n calls of the per-call CUDA-event time on this rank (ms)"""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        barrier()
        ev[0].record()
        for i in range(n):
            fn()
            ev[i + 1].record()
        barrier()
        return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(n)]))

    def exact_rows_check(ps, ci, rp, r0, Jr, S, seed, what):
        """sampled parity inside the bench: 64 random rows against exact integer arithmetic"""
        rng = np.random.default_rng(rank)
        rows = np.sort(rng.choice(Jr, size=min(64, Jr), replace=False))
        need = sorted(set(rows.tolist()) | {int(c) for r in rows for c in ci[rp[r]:rp[r + 1]]})
        host_rows = {r: synth.counts_host(seed, 0, 1, S, rows=[r0 + r], ld_cols=S)[0].astype(np.int64) for r in need}
        got = ps[torch.from_numpy(rows).to(dev)].cpu().numpy()
        for k, r in enumerate(rows):
            inc = host_rows[int(r)]
            exc = sum((host_rows[int(c)] for c in ci[rp[r]:rp[r + 1]]), np.zeros(S, dtype=np.int64))
            with np.errstate(divide="ignore", invalid="ignore"):
                want = (inc.astype(np.float32) / (inc.astype(np.float32) + exc.astype(np.float64))).astype(np.float32)
            if not np.array_equal(np.nan_to_num(got[k], nan=-1.0), np.nan_to_num(want, nan=-1.0)):
                raise SystemExit(f"rank {rank}: {what}: PS mismatch against exact arithmetic on row {r0 + int(r)}")

    S = args.samples
    J_total = args.junctions * world
    peak, peak_src = measured_peaks()
    launches = 0

    # ---- the job: junction set -> device cluster build -> row slabs --------------------------
    arrays = synth.junction_arrays(J_total, SEED)[:4]
    d_arrays = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev) for a in arrays]
    torch.cuda.synchronize()
    t_cold = time.perf_counter()
    cl = ops.cluster_build(*d_arrays)
    torch.cuda.synchronize()
    t_cold = time.perf_counter() - t_cold
    warm = []
    for _ in range(5):
        t0 = time.perf_counter()
        cl = ops.cluster_build(*d_arrays)
        torch.cuda.synchronize()
        warm.append(time.perf_counter() - t0)
    cluster_build = {"warm_ms": 1e3 * float(np.median(warm)), "cold_first_call_ms": t_cold * 1e3,
                     "junctions": J_total, "nnz": int(cl["nnz"]), "components": int(cl["n_comp"]),
                     "what": "sd_cluster_build + sd_cluster_fill on device-resident coordinates, host wall clock around the "
                             "call incl. its small readbacks; warm = median of 5 calls after the first (which also "
                             "pays workspace allocation and the sort library's first-use initialisation)"}
    del d_arrays
    row_ptr = cl["row_ptr"].cpu().numpy()
    col_idx = cl["col_idx"].cpu().numpy()
    parts = sharding.partition_rows(row_ptr, col_idx, world, sharding.row_weights(row_ptr, S))
    r0, r1 = parts[rank]
    rp, ci = sharding.shard_csr(row_ptr, col_idx, r0, r1)
    Jr = r1 - r0
    d_rp = torch.from_numpy(rp).to(dev)
    d_ci = torch.from_numpy(ci).to(dev)
    counts = ops.synth_counts(SEED + 1, r0, Jr, S, logical_cols=S, device=dev)      # rows keyed by GLOBAL row
    ps = torch.empty((Jr, S), dtype=torch.float32, device=dev)
    cells_rank = Jr * S

    def step():
        ops.quant_ps(counts, d_rp, d_ci, out_f32=ps, flags=args.flags)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    kernel_name = native.load().sd_quant_last_launch().decode()
    barrier()
    kernel_ms, graphed = timed_graph(step, args.steps, "the headline steps")
    launches += args.steps
    total_ms = max_over_ranks(kernel_ms * args.steps)
    cells_total = sum_over_ranks(float(cells_rank))
    ms_per_step = total_ms / args.steps
    value = cells_total / (ms_per_step * 1e-3)
    achieved = cells_rank * 8.0 / (kernel_ms * 1e-3) / 1e9
    exact_rows_check(ps, ci, rp, r0, Jr, S, SEED + 1, "headline")

    # ---- end to end: pinned host buffers through the host-pointer C-ABI call -------------------
    e2e = None
    h_ps = None
    if not args.no_e2e:
        h_counts = torch.empty((Jr, S), dtype=torch.int32).pin_memory()
        h_counts.copy_(counts)
        h_ps = torch.empty((Jr, S), dtype=torch.float32).pin_memory()
        n_e2e = max(3, min(args.steps, 7))
        for _ in range(5):       # warm-up: allocations, pool; the library times two calls per link format and keeps the faster
            ops.quant_ps_host(h_counts, rp, ci, out=h_ps, device=local_rank)
        per_call = []
        for _ in range(n_e2e):
            barrier()
            t0 = time.perf_counter()
            ops.quant_ps_host(h_counts, rp, ci, out=h_ps, device=local_rank)
            per_call.append(max_over_ranks(time.perf_counter() - t0))
        dt = float(np.mean(per_call))
        if not torch.equal(h_ps.view(torch.int32), ps.cpu().view(torch.int32)):
            raise SystemExit(f"rank {rank}: host-pipeline PS differs from the device-resident PS")
        h2d = int(Jr * S * 4 + rp.nbytes + ci.nbytes)
        d2h = int(Jr * S * 4)
        scratch = torch.empty(Jr * S * 4, dtype=torch.uint8, device=dev)
        h_sink = torch.empty(Jr * S * 4, dtype=torch.uint8).pin_memory()
        fabric = fabric_probe(dev, Jr * S * 4, d2h, chunk_mb=32, world=world,
                              buffers=(h_counts.view(torch.uint8).reshape(-1), h_sink, scratch,
                                       ps.view(torch.uint8).reshape(-1)))
        del scratch, h_sink
        e2e = {"value": cells_total / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": n_e2e,
               "ms_per_call": [round(t * 1e3, 3) for t in per_call],
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "api": "sd_quant_ps_host (pinned host counts -> pinned host PS), one call per step, wall clock, max over ranks",
               "fabric": fabric, "frac_of_fabric": fabric["duplex_ms"] / (dt * 1e3),
               "frac_of_fabric_note": "duplex_ms of the same bytes with no kernel / pipeline ms: 1.0 = the box's "
                                      "full-duplex pinned-copy ceiling; D2H alone is the lower bound of any pipeline",
               "host_cpus_local_to_gpu": len(numa_cpus) if numa_cpus else None}
        del h_counts

    # ---- the other epilogues of the same kernel on the same slab: counts_to_ps's float64 PS and
    # ir_table's intron-retention ratio (BASELINE.json configs[4]) --------------------------------
    variants = None
    if not args.no_variants:
        variants = {}
        try:
            med = torch.rand((Jr, S), dtype=torch.float64, device=dev).mul_(10).floor_()
            out64 = torch.empty((Jr, S), dtype=torch.float64, device=dev)

            def run_ir():
                native.call("sd_ir_ratio", Jr, S, native.ptr(med), med.stride(0), native.ptr(counts), counts.stride(0),
                            native.ptr(d_rp), native.ptr(d_ci), native.ptr(out64), out64.stride(0), 0, Jr,
                            native.stream_ptr())

            def run_f64():
                ops.quant_ps(counts, d_rp, d_ci, want_f32=False, out_f64=out64)

            for key, fn, bpc, what in (
                    ("ir_ratio", run_ir, 20, "sd_ir_ratio: median f64 + counts i32 -> IR f64 (ir_table.py:118-132; configs[4])"),
                    ("ps_f64", run_f64, 12, "sd_quant_ps float64 PS (counts_to_ps.py:62-68)")):
                for _ in range(2):
                    fn()
                name = native.load().sd_quant_last_launch().decode()
                v_ms, _ = timed_graph(fn, 5, key)
                launches += 5
                v_ms_max = max_over_ranks(v_ms)
                gbs = cells_rank * bpc / (v_ms * 1e-3) / 1e9
                variants[key] = {"ms_per_step": v_ms_max, "value": cells_total / (v_ms_max * 1e-3), "unit": UNIT,
                                 "bytes_per_cell": bpc, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak, "what": what,
                                 "kernel": name}
            del med, out64
        except torch.OutOfMemoryError:
            variants = None

    # ---- pairwise Fisher (configs[2]) ----------------------------------------------------------
    fisher = None
    Jf, Sf = args.fisher_junctions, args.fisher_samples
    pa, pb = ops.all_pairs(Sf)
    d_pa = torch.from_numpy(pa).to(dev); d_pb = torch.from_numpy(pb).to(dev)
    P = len(pa)
    if not args.no_fisher:
        f_arrays = synth.junction_arrays(Jf * world, SEED + 7)[:4]
        fcl = ops.cluster_build(*f_arrays)
        f_rp_all = fcl["row_ptr"].cpu().numpy(); f_ci_all = fcl["col_idx"].cpu().numpy()
        fparts = sharding.partition_rows(f_rp_all, f_ci_all, world)
        f0, f1 = fparts[rank]
        frp, fci = sharding.shard_csr(f_rp_all, f_ci_all, f0, f1)
        Jfr = f1 - f0
        # negative_binomial(2, p) = sum of two geometric draws
        inc = ops.synth_counts(SEED + 8, f0, Jfr, Sf, logical_cols=Sf, device=dev) + \
            ops.synth_counts(SEED + 9, f0, Jfr, Sf, logical_cols=Sf, device=dev)
        exc = ops.quant_ps(inc, frp, fci, want_f32=False, want_exc=True)["exc"]
        pout = torch.empty((Jfr, P), dtype=torch.float64, device=dev)
        # sd_fisher_pairwise_bounded: the bound on inc + exc is known, so each call is one asynchronous
        # kernel launch; every call is timed and the median reported (the shared hosts stall)
        n_f = max(5, min(args.steps, 9))
        bound = int((inc.long() + exc).max())           # measured once, outside the timed calls
        ops.fisher_pairwise(inc, exc, d_pa, d_pb, out=pout, max_cell_bound=bound)
        f_ms_rank = timed_calls(lambda: ops.fisher_pairwise(inc, exc, d_pa, d_pb, out=pout, max_cell_bound=bound), n_f)
        launches += n_f
        f_ms = max_over_ranks(f_ms_rank)
        tests_total = sum_over_ranks(float(Jfr * P))
        # work actually done per test: tail terms summed by the kernel's rule
        # on a sample of rows, 8 flop per term (4 DADD + 2 DMUL + 1 DFMA) + 250 flop of per-table setup;
        # SURVEY.md 8d's model (32 flop per support point) is reported beside it
        sel = np.sort(np.random.default_rng(5).choice(Jfr, size=min(120, Jfr), replace=False))
        inc_h = inc[torch.from_numpy(sel).to(dev)].cpu().numpy().astype(np.float64)
        exc_h = exc[torch.from_numpy(sel).to(dev)].cpu().numpy().astype(np.float64)
        cut_bits = 48
        terms, support, trivial = fisher_work(inc_h[:, pa], inc_h[:, pb], exc_h[:, pa], exc_h[:, pb], cut_bits)
        fp64_peak = ops.probe_fp64(dev)
        flops_per_test = 8.0 * terms + 250.0 * (1.0 - trivial)
        tests_per_s = tests_total / (f_ms * 1e-3)
        useful = flops_per_test * tests_per_s / world / 1e12
        cap = committed_capture("fisher_fp64.json") or {}
        fisher = {"metric": "fisher_tests_per_s", "value": tests_per_s, "unit": "tests/s",
                  "ms_per_step": f_ms, "steps": n_f, "timing": "median of per-call CUDA-event times, max over ranks",
                  "config": {"workload": f"pairwise Fisher: {Sf} samples ({P} pairs) x {Jf} junctions per GPU "
                                         f"(configs[2])", "mean_support": support, "trivial_fraction": trivial,
                             "mean_tail_terms_summed": terms, "tail_cut": f"2^-{cut_bits} of the running sum"},
                  "roofline": {"bound": "fp64", "achieved": useful, "peak": fp64_peak / 1e3, "unit": "TFLOP/s",
                               "frac": useful / (fp64_peak / 1e3), "flops_per_test": flops_per_test,
                               "model": "8 flop per summed tail term + 250 flop per non-trivial table (per GPU); the FP64 "
                                        "pipe issues DADD/DMUL at the DFMA rate and divergent lanes idle, so pipe "
                                        "occupancy (ncu) is the utilisation figure",
                               "pipe_fp64_active_ncu": ((cap.get("pipe_fp64_active_pct") or 0) / 100.0) or None,
                               "pipe_fp64_active_source": cap.get("source"),
                               "survey_model_tflops": 32.0 * support * tests_per_s / world / 1e12,
                               "peak_source": "sd_probe_fp64 (FMA microbenchmark, same run)"},
                  "dtype": "f64", "gpu_launches": n_f}
        # Benjamini-Hochberg per column on the same matrix (the CLI's default correction), device-resident
        try:
            padj = torch.empty_like(pout)
            ops.bh_adjust(pout, "pairwise", out=padj)
            b_ms = max_over_ranks(timed_calls(lambda: ops.bh_adjust(pout, "pairwise", out=padj), 3))
            fisher["bh"] = {"ms_per_step": b_ms, "value": tests_total / (b_ms * 1e-3), "unit": "p-values/s",
                            "what": "sd_bh_adjust, one Benjamini-Hochberg adjustment per sample pair (column) "
                                    "of the p-value matrix (pairwise_fisher.py:186-191)"}
            del padj
        except torch.OutOfMemoryError:
            fisher["bh"] = None
        # host-buffer API for the e2e figure (p-values come back to the host)
        if not args.no_e2e:
            inc_h_all = torch.empty(inc.shape, dtype=torch.int32).pin_memory(); inc_h_all.copy_(inc)
            out_h = torch.empty((Jfr, P), dtype=torch.float64).pin_memory()
            ops.pairwise_host(inc_h_all, frp, fci, pa, pb, out=out_h, device=local_rank)
            per_call = []
            for _ in range(3):
                barrier()
                t0 = time.perf_counter()
                ops.pairwise_host(inc_h_all, frp, fci, pa, pb, out=out_h, device=local_rank)
                per_call.append(max_over_ranks(time.perf_counter() - t0))
            dt = float(np.mean(per_call))
            if not torch.equal(out_h.view(torch.int64), pout.cpu().view(torch.int64)):
                raise SystemExit(f"rank {rank}: host-pipeline p-values differ from the device-resident ones")
            d2h_floor = None
            if e2e is not None:
                d2h_floor = out_h.numel() * 8 * world / (e2e["fabric"]["d2h_gbs"] * 1e9) * 1e3
            fisher["e2e"] = {"value": tests_total / dt, "unit": "tests/s", "ms_per_step": dt * 1e3,
                             "api": "sd_pairwise_host (pinned host inclusion counts + cluster CSR -> pinned host p-values; "
                                    "exclusion counts summed on the device)",
                             "h2d_bytes_per_step": int(inc_h_all.numel() * 4 + frp.nbytes + fci.nbytes),
                             "d2h_bytes_per_step": int(out_h.numel() * 8),
                             "d2h_alone_ms_at_fabric_rate": d2h_floor,
                             "frac_of_fabric": (d2h_floor / (dt * 1e3)) if d2h_floor else None}
            del out_h, inc_h_all
        del pout, inc, exc
        torch.cuda.empty_cache()

    # ---- strong scaling of the FIXED configs + the product's collectives (N > 1) -----------------
    strong = None
    collectives = None
    if world > 1 and not args.no_strong:
        strong, collectives = {}, {}
        nvlink_gbs = 900.0
        # configs[1] as ONE 400k x 1,000 problem (the N = 1 workload, same seed) cut over the N GPUs
        J1 = args.junctions
        a1 = synth.junction_arrays(J1, SEED)[:4]
        cl1 = ops.cluster_build(*a1)
        rp1, ci1 = cl1["row_ptr"].cpu().numpy(), cl1["col_idx"].cpu().numpy()
        parts1 = sharding.partition_rows(rp1, ci1, world, sharding.row_weights(rp1, S))
        q0, q1 = parts1[rank]
        srp, sci = sharding.shard_csr(rp1, ci1, q0, q1)
        full_counts = ops.synth_counts(SEED + 1, 0, J1, S, logical_cols=S, device=dev)
        full_ps = torch.empty((J1, S), dtype=torch.float32, device=dev)
        d_rp1, d_ci1 = cl1["row_ptr"], cl1["col_idx"]
        d_srp, d_sci = torch.from_numpy(srp).to(dev), torch.from_numpy(sci).to(dev)
        slab_counts = full_counts[q0:q1]
        slab_ps = torch.empty((q1 - q0, S), dtype=torch.float32, device=dev)
        tiny_rows = min(48, q1 - q0)

        def full_step():
            ops.quant_ps(full_counts, d_rp1, d_ci1, out_f32=full_ps, flags=args.flags)

        def slab_step():
            ops.quant_ps(slab_counts, d_srp, d_sci, out_f32=slab_ps, flags=args.flags)

        def floor_step():
            ops.quant_ps(slab_counts, d_srp, d_sci, out_f32=slab_ps, row_begin=0, row_end=tiny_rows, flags=args.flags)

        for fn in (full_step, floor_step, slab_step):
            for _ in range(3):
                fn()
        slab_name = native.load().sd_quant_last_launch().decode()
        t1_ms, _ = timed_graph(full_step, args.steps, "strong: full problem on one GPU")
        floor_ms, _ = timed_graph(floor_step, args.steps, "strong: launch floor")
        tn_ms, g_ok = timed_graph(slab_step, args.steps, "strong: row slab")
        launches += 3 * args.steps
        tn_max = max_over_ranks(tn_ms)
        t1_max = max_over_ranks(t1_ms)
        same = torch.equal(slab_ps.view(torch.int32), full_ps[q0:q1].view(torch.int32))
        strong["quant"] = {
            "workload": f"configs[1] as ONE {J1} x {S} problem cut into {world} row slabs at cluster boundaries "
                        f"(sharding.partition_rows), counts resident in HBM",
            "ms_per_step": tn_max, "ms_per_step_this_rank": tn_ms, "value": J1 * S / (tn_max * 1e-3), "unit": UNIT,
            "one_gpu_same_run_ms": t1_max, "speedup_vs_one_gpu": t1_max / tn_max,
            "launch_floor_ms": max_over_ranks(floor_ms),
            "launch_floor_note": "the same kernel on one 48-row tile inside the same kind of CUDA graph: what a launch "
                                 "costs before any row is processed",
            "slabs": parts1, "achieved_gbs_per_gpu": (q1 - q0) * S * 8 / (tn_ms * 1e-3) / 1e9,
            "slab_bits_equal_one_gpu": all_true(same), "launch_mode": "cuda graph" if g_ok else "stream launches",
            "kernel": slab_name, "scaling": "strong"}
        # collective 1: all-gather of the PS row slabs -> the whole matrix on every GPU
        gathered = sd_dist.all_gather_rows(slab_ps, parts1)
        ag_ms = max_over_ranks(timed_calls(lambda: sd_dist.all_gather_rows(slab_ps, parts1), 5))
        ag_same = torch.equal(gathered.view(torch.int32), full_ps.view(torch.int32))
        total_bytes = J1 * S * 4
        collectives["ps_all_gather"] = {
            "what": "distributed.all_gather_rows: NCCL all_gather of the float32 PS row slabs straight into the full "
                    "matrix (the optional device-resident full PS on every GPU)", "ms": ag_ms,
            "bytes_gathered_per_gpu": total_bytes, "algbw_gbs": total_bytes / (ag_ms * 1e-3) / 1e9,
            "busbw_gbs": total_bytes * (world - 1) / world / (ag_ms * 1e-3) / 1e9,
            "nvlink5_per_direction_gbs": nvlink_gbs,
            "frac_of_nvlink": total_bytes * (world - 1) / world / (ag_ms * 1e-3) / 1e9 / nvlink_gbs,
            "bits_equal_one_gpu_matrix": all_true(ag_same)}
        del gathered, full_counts, full_ps, slab_ps, slab_counts
        torch.cuda.empty_cache()

        # configs[2] as ONE 200k x 64 problem cut over the N GPUs (rows are independent once exc exists)
        if not args.no_fisher:
            a2 = synth.junction_arrays(Jf, SEED + 7)[:4]
            cl2 = ops.cluster_build(*a2)
            inc2 = ops.synth_counts(SEED + 8, 0, Jf, Sf, logical_cols=Sf, device=dev) + \
                ops.synth_counts(SEED + 9, 0, Jf, Sf, logical_cols=Sf, device=dev)
            exc2 = ops.quant_ps(inc2, cl2["row_ptr"], cl2["col_idx"], want_f32=False, want_exc=True)["exc"]
            bound2 = int((inc2.long() + exc2).max())
            cuts = [Jf * k // world for k in range(world + 1)]
            parts2 = [(cuts[k], cuts[k + 1]) for k in range(world)]
            g0, g1 = parts2[rank]
            p_full = torch.empty((Jf, P), dtype=torch.float64, device=dev)
            p_slab = torch.empty((g1 - g0, P), dtype=torch.float64, device=dev)
            inc_s, exc_s = inc2[g0:g1], exc2[g0:g1]
            ops.fisher_pairwise(inc2, exc2, d_pa, d_pb, out=p_full, max_cell_bound=bound2)
            ops.fisher_pairwise(inc_s, exc_s, d_pa, d_pb, out=p_slab, max_cell_bound=bound2)
            n_f = 5
            t1f = max_over_ranks(timed_calls(lambda: ops.fisher_pairwise(inc2, exc2, d_pa, d_pb, out=p_full, max_cell_bound=bound2), n_f))
            tnf_rank = timed_calls(lambda: ops.fisher_pairwise(inc_s, exc_s, d_pa, d_pb, out=p_slab, max_cell_bound=bound2), n_f)
            tnf = max_over_ranks(tnf_rank)
            launches += 2 * n_f
            same_f = torch.equal(p_slab.view(torch.int64), p_full[g0:g1].view(torch.int64))
            strong["fisher"] = {
                "workload": f"configs[2] as ONE {Jf} x {Sf} problem ({P} pairs) cut into {world} row slabs",
                "ms_per_step": tnf, "ms_per_step_this_rank": tnf_rank, "value": Jf * P / (tnf * 1e-3), "unit": "tests/s",
                "one_gpu_same_run_ms": t1f, "speedup_vs_one_gpu": t1f / tnf,
                "slab_bits_equal_one_gpu": all_true(same_f), "scaling": "strong",
                "timing": "median of per-call CUDA-event times, max over ranks"}
            # collective 2: the per-pair Benjamini-Hochberg correction ranks whole COLUMNS: all-to-all row
            # slabs -> column blocks, sd_bh_adjust on [J, P / N], all-to-all back
            adj = sd_dist.bh_columns_sharded(p_slab, parts2, lambda c: ops.bh_adjust(c, "pairwise", out=c))
            cols = sd_dist.rows_to_columns(p_slab, parts2)
            x1 = max_over_ranks(timed_calls(lambda: sd_dist.rows_to_columns(p_slab, parts2), 3))
            work = cols.clone()
            bh_ms = max_over_ranks(timed_calls(lambda: ops.bh_adjust(cols, "pairwise", out=work), 3))
            x2 = max_over_ranks(timed_calls(lambda: sd_dist.columns_to_rows(work, parts2, P), 3))
            whole = max_over_ranks(timed_calls(
                lambda: sd_dist.bh_columns_sharded(p_slab, parts2, lambda c: ops.bh_adjust(c, "pairwise", out=c)), 3))
            del cols, work
            ops.bh_adjust(p_full, "pairwise", out=p_full)               # single-GPU correction of the whole matrix
            same_adj = torch.equal(adj.view(torch.int64), p_full[g0:g1].view(torch.int64))
            # the same step with the exchange fused into the kernels: the Fisher kernel stores every p-value
            # into the column owner's matrix over NVLink peer memory (sd_fisher_pairwise_scatter), BH in place,
            # strided peer copies back -- no all-to-all, no packing (distributed.pairwise_fused)
            fused = None
            if world <= 16:
                def nccl_path():
                    ops.fisher_pairwise(inc_s, exc_s, d_pa, d_pb, out=p_slab, max_cell_bound=bound2)
                    return sd_dist.bh_columns_sharded(p_slab, parts2, lambda c: ops.bh_adjust(c, "pairwise", out=c))

                def fused_path():
                    return sd_dist.pairwise_fused(inc_s, exc_s, d_pa, d_pb, parts2, bound2)

                got_fused = fused_path()
                same_fused = torch.equal(got_fused.view(torch.int64), p_full[g0:g1].view(torch.int64))
                t_fused = max_over_ranks(timed_calls(fused_path, 5))
                t_nccl = max_over_ranks(timed_calls(nccl_path, 5))
                launches += 10
                fused = {"what": "distributed.pairwise_fused: Fisher with peer-memory scatter stores + barrier + sd_bh_adjust in "
                                 "place + one kernel of peer stores back, against Fisher + NCCL all_to_all + sd_bh_adjust + all_to_all "
                                 "(both: the whole per-slab pairwise step incl. Fisher)",
                         "fused_whole_ms": t_fused, "nccl_whole_ms": t_nccl, "speedup": t_nccl / t_fused,
                         "rows_bits_equal_one_gpu": all_true(same_fused)}
            slab_bytes = (g1 - g0) * P * 8
            sent = slab_bytes * (world - 1) / world
            collectives["pairwise_bh_exchange"] = {
                "what": "distributed.bh_columns_sharded: NCCL all_to_all (row slabs -> column blocks), sd_bh_adjust on "
                        "[J, P/N] columns, NCCL all_to_all back (pairwise_fisher.py:186-191 across ranks)",
                "rows_to_columns_ms": x1, "bh_adjust_ms": bh_ms, "columns_to_rows_ms": x2, "whole_ms": whole,
                "bytes_sent_per_gpu_each_way": int(sent),
                "all_to_all_gbs_per_gpu_each_way": [sent / (x1 * 1e-3) / 1e9, sent / (x2 * 1e-3) / 1e9],
                "nvlink5_per_direction_gbs": nvlink_gbs,
                "adjusted_rows_bits_equal_one_gpu": all_true(same_adj), "fused": fused,
                "note": "exchange times include the column-block packing (contiguous copies) and torch.cat on arrival"}
            del adj, p_full, p_slab, inc2, exc2
            torch.cuda.empty_cache()

    # ---- configs[3]: 10,000 samples x 1M junctions cut over the N GPUs ----------------------------
    tcga = None
    if not args.no_tcga:
        J3, S3 = args.tcga_junctions, args.tcga_samples
        try:
            del counts, ps
            torch.cuda.empty_cache()
            a3 = synth.junction_arrays(J3, SEED + 3)[:4]
            cl3 = ops.cluster_build(*a3)
            rp3_all, ci3_all = cl3["row_ptr"].cpu().numpy(), cl3["col_idx"].cpu().numpy()
            parts3 = sharding.partition_rows(rp3_all, ci3_all, world, sharding.row_weights(rp3_all, S3))
            t0_, t1_ = parts3[rank]
            rp3, ci3 = sharding.shard_csr(rp3_all, ci3_all, t0_, t1_)
            d_rp3, d_ci3 = torch.from_numpy(rp3).to(dev), torch.from_numpy(ci3).to(dev)
            c3 = ops.synth_counts(SEED + 4, t0_, t1_ - t0_, S3, logical_cols=S3, device=dev)
            ps3 = torch.empty((t1_ - t0_, S3), dtype=torch.float32, device=dev)

            def step3():
                ops.quant_ps(c3, d_rp3, d_ci3, out_f32=ps3)

            for _ in range(2):
                step3()
            name3 = native.load().sd_quant_last_launch().decode()
            k3 = 5
            ms3_rank, _ = timed_graph(step3, k3, "configs[3]")
            launches += k3
            ms3 = max_over_ranks(ms3_rank)
            exact_rows_check(ps3, ci3, rp3, t0_, t1_ - t0_, S3, SEED + 4, "configs[3]")
            gbs3 = (t1_ - t0_) * S3 * 8 / (ms3_rank * 1e-3) / 1e9
            tcga = {"workload": f"BASELINE.json configs[3]: quant PS, {S3} samples x {J3} junctions, cluster-sharded over "
                                f"{world} GPU(s) ({(t1_ - t0_) * S3 * 8 / 1e9:.1f} GB of counts + PS on this rank)",
                    "ms_per_step": ms3, "value": J3 * S3 / (ms3 * 1e-3), "unit": UNIT, "steps": k3,
                    "achieved_gbs_per_gpu": gbs3, "frac_of_hbm_peak": gbs3 / peak, "slabs": parts3,
                    "nnz": int(rp3_all[-1]), "sampled_rows_exact": True, "kernel": name3, "scaling": "strong"}
            del c3, ps3
            torch.cuda.empty_cache()
        except torch.OutOfMemoryError:
            tcga = {"skipped": "out of device memory"}

    clocks = sampler.stop()
    clocks["window"] = "warm-up, timed steps, e2e, Fisher, strong-scaling and configs[3] sections (the timed K2 region alone lasts ~10 ms)"

    # ---- CPU baseline (rank 0, N = 1 only): the unmodified reference on the full workload ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        workers = args.cpu_workers or cores
        rows_sorted, c32 = host_workload(args.junctions, S)
        fan = CpuQuant(rows_sorted, c32, workers)
        kind, workers = fan.kind, fan.workers
        fan.step()                                                   # warm-up (page faults of the workers' first pass)
        dt_all = float(np.mean([fan.step() for _ in range(2)]))
        fan.close()
        keep = np.zeros((args.junctions, S), dtype=np.float32)
        one = CpuQuant(rows_sorted, c32, 1, keep=keep)
        dt_one = one.step()
        one.close()
        parity = None
        if h_ps is not None and tuple(h_ps.shape) == keep.shape:
            parity = bool(np.array_equal(h_ps.numpy().view(np.uint32), keep.view(np.uint32)))
            if not parity:
                raise SystemExit("the reference's PS matrix differs from the GPU's (full size, bit for bit)")
        cpu = {"value": args.junctions * S / dt_all, "unit": UNIT, "cores": workers, "kind": kind, "host_cores": cores,
               "sample": reference_sample_text(args, workers, kind) + f", {dt_all:.2f} s",
               "one_core": {"value": args.junctions * S / dt_one, "seconds": dt_one,
                            "what": "the same three reference stages in one process on the whole problem (the reference "
                                    "is single-threaded)"},
               "ps_bits_equal_gpu_full_size": parity}
        del keep, c32
        if fisher is not None:
            tps, fkind = cpu_fisher_tests_per_s(args.cpu_fisher_events, args.fisher_samples)
            fisher["cpu_baseline"] = {"value": tps, "unit": "tests/s", "cores": 1, "kind": fkind,
                                      "sample": f"{args.cpu_fisher_events} events x "
                                                f"{args.fisher_samples * (args.fisher_samples - 1) // 2} pairs through "
                                                + ("the unmodified pairwise_fisher.run_with (oracle/_ref; correction none; "
                                                   "scipy.stats.fisher_exact per table)" if fkind == "reference" else
                                                   "oracle/ref_port.py pairwise_loop")}

    traffic_cap = committed_capture("quant_traffic.json") or {}
    traffic = None
    if traffic_cap and traffic_cap.get("samples") == S and abs(traffic_cap.get("rows", 0) - Jr) <= 0.01 * Jr:
        traffic = int((traffic_cap["dram_bytes_read"] + traffic_cap["dram_bytes_write"]) * Jr / traffic_cap["rows"])
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "i32/i64 sums, f32 divide", "data": "synthetic",
            "config": base_config(args),
            "job": {"junctions_total": J_total, "nnz": int(row_ptr[-1]), "slabs": parts, "flags": args.flags,
                    "l2": f"inputs+outputs {cells_rank * 8 / 1e9:.1f} GB per pass per GPU >> 126 MB L2, no flush"},
            "cluster_build": cluster_build,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_cap.get("source") if traffic else None,
                         "traffic_note": ("ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at "
                                          "this shape, from the committed capture named in traffic_source (a profiler cannot "
                                          "run inside the timed bench)") if traffic else None,
                         "kernel": kernel_name, "kernel_ms": kernel_ms,
                         "bytes_per_cell": 8, "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps,
            "gpu_launches_all_timed_regions": launches,
            "gpu_launches_note": f"{args.steps} quant kernel launches in the headline timed region (one per step); "
                                 f"{launches} counting the variant / Fisher / strong-scaling / configs[3] timed regions too",
            "launch_mode": "cuda graph of K kernel launches" if graphed else "K stream launches",
            "clocks": clocks, "variants": variants, "fisher": fisher, "strong": strong, "collectives": collectives,
            "tcga": tcga,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
