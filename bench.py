#!/usr/bin/env python
"""Benchmark of the SpliceDICE quant / pairwise hot path on B200 (one JSON line on stdout).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1

Workload (BASELINE.json configs[1]): quant PS, 1,000 samples x 400,000 junctions per GPU
(weak scaling: N GPUs hold N contiguous row slabs of an N x 400k-junction problem, cut where
no adjacency edge crosses -- splicedice_b200/sharding.py).  A step is one pass of the fused
exclusion-aggregation + PS kernel over the rank's slab.

  value        PS cells/s, counts and PS resident in HBM (CUDA events, max over ranks)
  e2e          the same metric through sd_quant_ps_host: pinned host counts in, pinned host PS
               out, H2D and D2H inside the timed region
  roofline     8 algorithmic bytes per cell (4 B int32 count read + 4 B float32 PS written)
               over the kernel's average launch time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline oracle/ref_port.py (a loop-for-loop port of SPLICEDICE.getClusters +
               calculatePsi) on a bounded row sample, one host core (the reference is
               single-threaded)
  fisher       pairwise Fisher (configs[2]: 64 samples, 2,016 pairs x 200,000 junctions):
               tests/s, FP64 model, its own CPU baseline

`--impl reference` times only the CPU port, fanned out over all host cores.
Inputs are larger than L2 (1.6 GB in + 1.6 GB out per pass against 126 MB), so no flush is
needed between timed iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
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
    ap.add_argument("--cpu-rows", type=int, default=150_000, help="rows of the CPU-baseline sample")
    ap.add_argument("--cpu-fisher-events", type=int, default=6)
    ap.add_argument("--flags", type=int, default=0, help="sd_quant_ps flags (kernel variant / tile shape)")
    return ap.parse_args()


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


def fisher_pipe_active():
    """sm__inst_executed_pipe_fp64 (% of peak) of the Fisher kernel from the committed ncu capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "fisher_fp64.json")))["pipe_fp64_active_pct"] / 100.0
    except Exception:
        return None


def fisher_work(a, b, c, d):
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
            done |= (P.view(np.int64) >> 32) < (A.view(np.int64) >> 32) - (48 << 20)
            if done.all():
                break
        return count
    with np.errstate(all="ignore"):
        near = bursts(a2, d2, b2, c2)
        g = np.minimum(2 * mode2 - a2, np.minimum(n1, nn))          # mirror point ~ first admitted far-side point
        far = bursts(n1 - g, nn - g, g, n2 - nn + g)
    terms = np.where(known, 0, 4 * (near + far))
    return float(terms.mean()), float(support.mean()), float(trivial.mean())


def config_label(samples, junctions_total):
    if samples == 1000 and junctions_total % 400_000 == 0:
        return "BASELINE.json configs[1], GTEx-scale"
    if samples == 10_000 and junctions_total == 1_000_000:
        return "BASELINE.json configs[3], TCGA-scale, cluster-sharded"
    return "custom shape"


def ncu_traffic(rows, samples):
    """dram__bytes_read + dram__bytes_write of one launch from the committed ncu capture
    (profiles/), valid only for the shape it was taken on."""
    path = os.path.join(ROOT, "profiles", "quant_traffic.json")
    try:
        d = json.load(open(path))
        if d["samples"] == samples and abs(d["rows"] - rows) <= 0.01 * d["rows"]:
            scale = rows / d["rows"]          # slabs differ from the captured shape by < 1 % of rows
            return int((d["dram_bytes_read"] + d["dram_bytes_write"]) * scale), d["source"]
    except Exception:
        pass
    return None, None


# ------------------------------------------------------------------------------------------
# CPU baseline (oracle port; also the --impl reference arm)
# ------------------------------------------------------------------------------------------
def cpu_quant_sample(n_rows, n_samples, seed):
    """Junction tuples + float32 counts of an n_rows sample of the workload (host generator)."""
    from splicedice_b200 import synth
    js = synth.junction_tuples(n_rows, seed)
    counts = synth.counts_host(seed + 1, 0, n_rows, n_samples).astype(np.float32)
    return js, counts


_FORK_STATE = {}


def _psi_slab(slab):
    """One worker's share of the psi loop; inputs are inherited through fork (no pickling), only a
    checksum travels back."""
    from oracle import ref_port
    adjacency, index, counts, keys, workers = (_FORK_STATE[k] for k in ("adjacency", "index", "counts", "keys", "workers"))
    out = ref_port.psi_rows(adjacency, index, counts, keys[slab::workers])
    return float(np.nansum(out))


def cpu_quant_time(js, counts, workers=1):
    """Seconds of the reference's getClusters + row index + calculatePsi (ported loops); with
    workers > 1 the psi rows are dealt round-robin to forked processes."""
    from oracle import ref_port
    t0 = time.perf_counter()
    adjacency = ref_port.sweep_clusters(js)
    index = ref_port.row_index(adjacency)
    if workers <= 1:
        ref_port.psi_loop(adjacency, index, counts)
    else:
        import multiprocessing as mp
        _FORK_STATE.update(adjacency=adjacency, index=index, counts=counts, keys=sorted(adjacency), workers=workers)
        with mp.get_context("fork").Pool(workers) as pool:
            pool.map(_psi_slab, range(workers))
        _FORK_STATE.clear()
    return time.perf_counter() - t0


def cpu_fisher_time(n_events, n_samples, seed):
    from oracle import ref_port
    rng = np.random.default_rng(seed)
    inc = rng.negative_binomial(2, 0.02, size=(n_events, n_samples)).astype(np.float64)
    names = np.array([f"chr1:{100 * i}-{100 * i + 150}:+" for i in range(n_events)])
    clusters = {names[i]: [names[j] for j in (i - 1, i + 1) if 0 <= j < n_events] for i in range(n_events)}
    t0 = time.perf_counter()
    ref_port.pairwise_loop(names, inc, clusters)
    dt = time.perf_counter() - t0
    return n_events * (n_samples * (n_samples - 1) // 2) / dt


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    js, counts = cpu_quant_sample(args.cpu_rows, args.samples, SEED)
    cells = args.cpu_rows * args.samples
    times = []
    for i in range(args.warmup + args.steps):
        dt = cpu_quant_time(js, counts, workers=cores)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = cells / (ms * 1e-3)
    sample = (f"{args.cpu_rows} junctions x {args.samples} samples of the workload per step: "
              f"sweep_clusters + row_index + psi_loop (oracle/ref_port.py), psi rows fanned over {cores} processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"quant PS: {args.samples} samples x {args.junctions} junctions per GPU (configs[1])",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    S = args.samples
    J_total = args.junctions * world

    # ---- the job: junction set -> device cluster build -> row slabs --------------------------
    arrays = synth.junction_arrays(J_total, SEED)[:4]
    t_k1 = time.perf_counter()
    cl = ops.cluster_build(*arrays)
    torch.cuda.synchronize()
    t_k1 = time.perf_counter() - t_k1
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
    barrier()
    # The K timed steps are captured into one CUDA graph, so a busy host cannot open gaps between
    # the launches (each step is still its own kernel launch inside the graph).  If capture is not
    # possible the steps are launched one by one.
    graph = None
    try:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(args.steps):
                    step()
        torch.cuda.current_stream().wait_stream(side)
        graph = g
    except Exception as exc_capture:            # noqa: BLE001 - any capture failure falls back to plain launches
        print(f"[bench] CUDA-graph capture unavailable ({exc_capture}); launching step by step", file=sys.stderr)
        torch.cuda.synchronize()
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_begin.record()
    if graph is not None:
        graph.replay()
    else:
        for _ in range(args.steps):
            step()
    e_end.record()
    barrier()
    total_ms_rank = e_begin.elapsed_time(e_end)
    per_launch_ms = [total_ms_rank / args.steps]
    total_ms = max_over_ranks(total_ms_rank)
    cells_total = sum_over_ranks(float(cells_rank))
    ms_per_step = total_ms / args.steps
    value = cells_total / (ms_per_step * 1e-3)
    kernel_ms = float(np.mean(per_launch_ms))
    peak, peak_src = measured_peaks()
    achieved = cells_rank * 8.0 / (kernel_ms * 1e-3) / 1e9

    # sampled parity inside the bench: 64 random rows against exact integer arithmetic
    rng = np.random.default_rng(rank)
    rows = np.sort(rng.choice(Jr, size=min(64, Jr), replace=False))
    need = sorted(set(rows.tolist()) | {int(c) for r in rows for c in ci[rp[r]:rp[r + 1]]})
    host_rows = {r: synth.counts_host(SEED + 1, 0, 1, S, rows=[r0 + r], ld_cols=S)[0].astype(np.int64) for r in need}
    got = ps[torch.from_numpy(rows).to(dev)].cpu().numpy()
    for k, r in enumerate(rows):
        inc = host_rows[int(r)]
        exc = sum((host_rows[int(c)] for c in ci[rp[r]:rp[r + 1]]), np.zeros(S, dtype=np.int64))
        with np.errstate(divide="ignore", invalid="ignore"):
            want = (inc.astype(np.float32) / (inc.astype(np.float32) + exc.astype(np.float64))).astype(np.float32)
        if not np.array_equal(np.nan_to_num(got[k], nan=-1.0), np.nan_to_num(want, nan=-1.0)):
            raise SystemExit(f"rank {rank}: PS mismatch against exact arithmetic on row {r0 + int(r)}")

    # ---- end to end: pinned host buffers through the host-pointer C-ABI call -------------------
    e2e = None
    if not args.no_e2e:
        h_counts = torch.empty((Jr, S), dtype=torch.int32).pin_memory()
        h_counts.copy_(counts)
        h_ps = torch.empty((Jr, S), dtype=torch.float32).pin_memory()
        n_e2e = max(2, min(args.steps, 5))
        ops.quant_ps_host(h_counts, rp, ci, out=h_ps, device=local_rank)          # warm-up (allocations, pool)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            ops.quant_ps_host(h_counts, rp, ci, out=h_ps, device=local_rank)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0) / n_e2e
        if not torch.equal(h_ps.view(torch.int32), ps.cpu().view(torch.int32)):
            raise SystemExit(f"rank {rank}: host-pipeline PS differs from the device-resident PS")
        e2e = {"value": cells_total / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": n_e2e,
               "h2d_bytes_per_step": int(Jr * S * 4 + rp.nbytes + ci.nbytes), "d2h_bytes_per_step": int(Jr * S * 4),
               "api": "sd_quant_ps_host (pinned host counts -> pinned host PS)",
               "host_cpus_local_to_gpu": len(numa_cpus) if numa_cpus else None}
        del h_counts, h_ps

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
                barrier()
                v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                v0.record()
                for _ in range(5):
                    fn()
                v1.record()
                barrier()
                v_ms = max_over_ranks(v0.elapsed_time(v1) / 5)
                gbs = cells_rank * bpc / (v_ms * 1e-3) / 1e9
                variants[key] = {"ms_per_step": v_ms, "value": cells_total / (v_ms * 1e-3), "unit": UNIT,
                                 "bytes_per_cell": bpc, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak, "what": what}
            del med, out64
        except torch.OutOfMemoryError:
            variants = None

    # ---- pairwise Fisher (configs[2]) ----------------------------------------------------------
    fisher = None
    if not args.no_fisher:
        Jf, Sf = args.fisher_junctions, args.fisher_samples
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
        pa, pb = ops.all_pairs(Sf)
        d_pa = torch.from_numpy(pa).to(dev); d_pb = torch.from_numpy(pb).to(dev)
        pout = torch.empty((Jfr, len(pa)), dtype=torch.float64, device=dev)
        # sd_fisher_pairwise_bounded: the bound on inc + exc is known, so each call is one asynchronous
        # kernel launch; every call is timed and the median reported (the shared hosts stall)
        n_f = max(5, min(args.steps, 9))
        bound = int((inc.long() + exc).max())           # measured once, outside the timed calls
        ops.fisher_pairwise(inc, exc, d_pa, d_pb, out=pout, max_cell_bound=bound)
        barrier()
        fev = [torch.cuda.Event(enable_timing=True) for _ in range(n_f + 1)]
        fev[0].record()
        for i in range(n_f):
            ops.fisher_pairwise(inc, exc, d_pa, d_pb, out=pout, max_cell_bound=bound)
            fev[i + 1].record()
        barrier()
        f_ms = max_over_ranks(float(np.median([fev[i].elapsed_time(fev[i + 1]) for i in range(n_f)])))
        tests_total = sum_over_ranks(float(Jfr * len(pa)))
        # work actually done per test: tail terms summed by the kernel's rule
        # (cut at 2^-48 of the running sum, checked every 4 terms) on a sample of rows, 8 flop per
        # term (4 DADD + 2 DMUL + 1 DFMA) + 250 flop of per-table setup; SURVEY.md 8d's model
        # (32 flop per support point) is reported beside it
        sel = np.sort(np.random.default_rng(5).choice(Jfr, size=min(120, Jfr), replace=False))
        inc_h = inc[torch.from_numpy(sel).to(dev)].cpu().numpy().astype(np.float64)
        exc_h = exc[torch.from_numpy(sel).to(dev)].cpu().numpy().astype(np.float64)
        terms, support, trivial = fisher_work(inc_h[:, pa], inc_h[:, pb], exc_h[:, pa], exc_h[:, pb])
        fp64_peak = ops.probe_fp64(dev)
        flops_per_test = 8.0 * terms + 250.0 * (1.0 - trivial)
        tests_per_s = tests_total / (f_ms * 1e-3)
        useful = flops_per_test * tests_per_s / world / 1e12
        fisher = {"metric": "fisher_tests_per_s", "value": tests_per_s, "unit": "tests/s",
                  "ms_per_step": f_ms, "steps": n_f, "timing": "median of per-call CUDA-event times, max over ranks",
                  "config": {"workload": f"pairwise Fisher: {Sf} samples ({len(pa)} pairs) x {Jf} junctions per GPU "
                                         f"(configs[2])", "mean_support": support, "trivial_fraction": trivial,
                             "mean_tail_terms_summed": terms},
                  "roofline": {"bound": "fp64", "achieved": useful, "peak": fp64_peak / 1e3, "unit": "TFLOP/s",
                               "frac": useful / (fp64_peak / 1e3), "flops_per_test": flops_per_test,
                               "model": "8 flop per summed tail term + 250 flop per non-trivial table (per GPU); the FP64 "
                                        "pipe issues DADD/DMUL at the DFMA rate and divergent lanes idle, so pipe "
                                        "occupancy (ncu) is the utilisation figure",
                               "pipe_fp64_active_ncu": fisher_pipe_active(),
                               "survey_model_tflops": 32.0 * support * tests_per_s / world / 1e12,
                               "peak_source": "sd_probe_fp64 (FMA microbenchmark, same run)"},
                  "dtype": "f64", "gpu_launches": n_f}
        # Benjamini-Hochberg per column on the same matrix (the CLI's default correction), device-resident
        try:
            padj = torch.empty_like(pout)
            ops.bh_adjust(pout, "pairwise", out=padj)
            barrier()
            bev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            bev[0].record()
            for i in range(3):
                ops.bh_adjust(pout, "pairwise", out=padj)
                bev[i + 1].record()
            barrier()
            b_ms = max_over_ranks(float(np.median([bev[i].elapsed_time(bev[i + 1]) for i in range(3)])))
            fisher["bh"] = {"ms_per_step": b_ms, "value": tests_total / (b_ms * 1e-3), "unit": "p-values/s",
                            "what": "sd_bh_adjust, one Benjamini-Hochberg adjustment per sample pair (column) "
                                    "of the p-value matrix (pairwise_fisher.py:186-191)"}
            del padj
        except torch.OutOfMemoryError:
            fisher["bh"] = None
        # host-buffer API for the e2e figure (p-values come back to the host)
        if not args.no_e2e:
            inc_h_all = torch.empty(inc.shape, dtype=torch.int32).pin_memory(); inc_h_all.copy_(inc)
            exc_h_all = torch.empty(exc.shape, dtype=torch.int64).pin_memory(); exc_h_all.copy_(exc)
            out_h = torch.empty((Jfr, len(pa)), dtype=torch.float64).pin_memory()
            ops.fisher_pairwise_host(inc_h_all, exc_h_all, pa, pb, out=out_h, device=local_rank)
            barrier()
            t0 = time.perf_counter()
            ops.fisher_pairwise_host(inc_h_all, exc_h_all, pa, pb, out=out_h, device=local_rank)
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            fisher["e2e"] = {"value": tests_total / dt, "unit": "tests/s", "ms_per_step": dt * 1e3,
                             "h2d_bytes_per_step": int(inc_h_all.numel() * 4 + exc_h_all.numel() * 8),
                             "d2h_bytes_per_step": int(out_h.numel() * 8)}
            del out_h
        del pout

    clocks = sampler.stop()
    clocks["window"] = "warm-up, timed steps, e2e and Fisher sections (the timed K2 region alone lasts ~10 ms)"

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        js, c32 = cpu_quant_sample(args.cpu_rows, S, SEED)
        dt = cpu_quant_time(js, c32, workers=1)
        cpu = {"value": args.cpu_rows * S / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "host_cores": os.cpu_count(),
               "sample": f"{args.cpu_rows} junctions x {S} samples: oracle/ref_port.py sweep_clusters + row_index + "
                         f"psi_loop (loop-for-loop port of SPLICEDICE.py:230-255,96,297-310), {dt:.1f} s, 1 core"}
        if fisher is not None:
            tps = cpu_fisher_time(args.cpu_fisher_events, args.fisher_samples, SEED)
            fisher["cpu_baseline"] = {"value": tps, "unit": "tests/s", "cores": 1, "kind": "port",
                                      "sample": f"{args.cpu_fisher_events} events x "
                                                f"{args.fisher_samples * (args.fisher_samples - 1) // 2} pairs: "
                                                f"oracle/ref_port.py pairwise_loop (scipy.stats.fisher_exact per table)"}

    traffic, traffic_src = ncu_traffic(Jr, S)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "i32/i64 sums, f32 divide", "data": "synthetic",
            "config": {"workload": f"quant PS: {S} samples x {args.junctions} junctions per GPU ({config_label(S, J_total)})",
                       "junctions_total": J_total, "nnz": int(row_ptr[-1]), "slabs": parts,
                       "l2": f"inputs+outputs {cells_rank * 8 / 1e9:.1f} GB per pass per GPU >> 126 MB L2, no flush", "seed": SEED,
                       "cluster_build_ms": t_k1 * 1e3, "flags": args.flags},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": "quant_wide_kernel<2, kOutF32> (256-column slabs, 48-row TMA tiles)", "kernel_ms": kernel_ms,
                         "bytes_per_cell": 8, "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps, "launch_mode": "cuda graph of K kernel launches" if graph is not None else "K stream launches",
            "clocks": clocks, "variants": variants, "fisher": fisher,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
