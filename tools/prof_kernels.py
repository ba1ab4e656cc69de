#!/usr/bin/env python
"""Small driver for ncu captures: runs the quant PS kernel (configs[1] shape, fewer rows) and/or
the pairwise Fisher kernel (configs[2] shape, fewer rows) a few times.

    python tools/prof_kernels.py quant [rows] [samples] [flags]
    python tools/prof_kernels.py fisher [rows] [samples]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from splicedice_b200 import ops, synth  # noqa: E402


def main():
    what = sys.argv[1]
    dev = torch.device("cuda", 0)
    if what == "quant":
        J = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
        S = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
        flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
        cl = ops.cluster_build(*synth.junction_arrays(J, 20261018)[:4])
        counts = ops.synth_counts(1, 0, J, S, device=dev)
        ps = torch.empty((J, S), dtype=torch.float32, device=dev)
        if os.environ.get("SD_PROF_EMPTY_CSR"):           # structural floor: no neighbours at all
            cl["row_ptr"].zero_()
        if os.environ.get("SD_PROF_INTILE_ONLY"):         # drop every neighbour outside its row's R-row tile
            R = int(os.environ["SD_PROF_INTILE_ONLY"])
            rp, ci = cl["row_ptr"].cpu().numpy(), cl["col_idx"].cpu().numpy()
            rows = np.repeat(np.arange(J), np.diff(rp))
            keep = (rows // R) == (ci // R)
            print(f"  in-tile entries: {keep.mean():.3f} of {len(ci)}")
            rp2 = np.zeros(J + 1, dtype=np.int32)
            np.cumsum(np.bincount(rows[keep], minlength=J), out=rp2[1:])
            cl["row_ptr"] = torch.from_numpy(rp2).to(dev)
            cl["col_idx"] = torch.from_numpy(ci[keep].astype(np.int32)).to(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(4):
            if i == 1:
                e0.record()
            ops.quant_ps(counts, cl["row_ptr"], cl["col_idx"], out_f32=ps, flags=flags)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"quant {J}x{S} flags={flags}: {ms:.3f} ms/launch, {J * S * 8 / ms / 1e6:.0f} GB/s algorithmic")
    else:
        J = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
        S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
        cl = ops.cluster_build(*synth.junction_arrays(J, 20261025)[:4])
        inc = ops.synth_counts(8, 0, J, S, device=dev) + ops.synth_counts(9, 0, J, S, device=dev)
        exc = ops.quant_ps(inc, cl["row_ptr"], cl["col_idx"], want_f32=False, want_exc=True)["exc"]
        pa, pb = ops.all_pairs(S)
        out = torch.empty((J, len(pa)), dtype=torch.float64, device=dev)
        reps = int(os.environ.get("SD_PROF_REPS", "2"))
        bound = int((inc.long() + exc).max())     # asynchronous form: kernel time without the host round trip
        pa, pb = torch.from_numpy(pa).to(dev), torch.from_numpy(pb).to(dev)
        for _ in range(int(os.environ.get("SD_PROF_WARM", "1"))):       # let the clocks settle before timing
            ops.fisher_pairwise(inc, exc, pa, pb, out=out, max_cell_bound=bound)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            ops.fisher_pairwise(inc, exc, pa, pb, out=out, max_cell_bound=bound)
            ev[i + 1].record()
        torch.cuda.synchronize()
        per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
        ms = per[len(per) // 2]
        print(f"  per-call ms: min {per[0]:.3f} median {ms:.3f} max {per[-1]:.3f}")
        print(f"fisher {J}x{len(pa)}: {ms:.3f} ms/launch, {J * len(pa) / (ms * 1e-3):.3e} tests/s")




def misc():
    """K1 timing, narrow-matrix quant (configs[0] shape), IR (configs[4]) -- CUDA events."""
    import time
    dev = torch.device("cuda", 0)
    for J in (50_000, 400_000, 1_000_000):
        arrays = synth.junction_arrays(J, 3)[:4]
        d = [torch.from_numpy(a).to(dev) for a in arrays]
        ops.cluster_build(*d)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            cl = ops.cluster_build(*d)
        torch.cuda.synchronize()
        print(f"cluster_build J={J}: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms/call (nnz {cl['nnz']}, {cl['n_comp']} components)")
    for J, S in ((50_000, 8), (400_000, 64), (400_000, 1000)):
        cl = ops.cluster_build(*synth.junction_arrays(J, 3)[:4])
        counts = ops.synth_counts(1, 0, J, S, device=dev)
        ps = torch.empty((J, S), dtype=torch.float32, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(12):
            if i == 2:
                e0.record()
            ops.quant_ps(counts, cl["row_ptr"], cl["col_idx"], out_f32=ps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"quant {J}x{S}: {ms * 1e3:.1f} us/launch, {J * S * 8 / ms / 1e6:.0f} GB/s algorithmic")
        if S == 1000:
            med = torch.rand((J, S), dtype=torch.float64, device=dev).mul_(10).floor_()
            for i in range(6):
                if i == 1:
                    e0.record()
                ir = ops.ir_ratio(med, counts, cl["row_ptr"], cl["col_idx"])
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"ir_ratio {J}x{S}: {ms:.3f} ms/launch (incl. output allocation), {J * S * 20 / ms / 1e6:.0f} GB/s algorithmic (20 B/cell)")
            ps64 = torch.empty((J, S), dtype=torch.float64, device=dev)
            for i in range(6):
                if i == 1:
                    e0.record()
                ops.quant_ps(counts, cl["row_ptr"], cl["col_idx"], want_f32=False, out_f64=ps64)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"quant f64 PS {J}x{S}: {ms:.3f} ms/launch, {J * S * 12 / ms / 1e6:.0f} GB/s algorithmic (12 B/cell)")


def variants():
    """configs[4] (ir_table ratio) and the float64 PS of counts_to_ps at 400,000 x 1,000: CUDA events."""
    dev = torch.device("cuda", 0)
    J = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    cl = ops.cluster_build(*synth.junction_arrays(J, 3)[:4])
    counts = ops.synth_counts(1, 0, J, S, device=dev)
    med = torch.rand((J, S), dtype=torch.float64, device=dev).mul_(10).floor_()
    ir = torch.empty((J, S), dtype=torch.float64, device=dev)
    ps64 = torch.empty((J, S), dtype=torch.float64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from splicedice_b200 import native
    for name, fn, bytes_per_cell in (
            ("ir_ratio", lambda: native.call("sd_ir_ratio", J, S, native.ptr(med), med.stride(0), native.ptr(counts),
                                             counts.stride(0), native.ptr(cl["row_ptr"]), native.ptr(cl["col_idx"]),
                                             native.ptr(ir), ir.stride(0), 0, J, native.stream_ptr()), 20),
            ("quant f64 PS", lambda: ops.quant_ps(counts, cl["row_ptr"], cl["col_idx"], want_f32=False, out_f64=ps64), 12)):
        for i in range(7):
            if i == 2:
                e0.record()
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{name} {J}x{S}: {ms:.3f} ms/launch, {J * S * bytes_per_cell / ms / 1e6:.0f} GB/s algorithmic ({bytes_per_cell} B/cell)")


if len(sys.argv) > 1 and sys.argv[1] == "misc":
    misc()
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "variants":
    variants()
    sys.exit(0)


if __name__ == "__main__":
    main()
