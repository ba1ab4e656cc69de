"""Time the device Benjamini-Hochberg (sd_bh_adjust) on the configs[2] p-value matrix
(200,000 junctions x 2,016 pairs) next to the numpy restatement on a sample of columns."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import oracle_np                      # noqa: E402  (CPU baseline only)
from splicedice_b200 import ops                   # noqa: E402


def main():
    J = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 2016
    ops.require_cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    p = torch.rand((J, P), dtype=torch.float64, device="cuda", generator=g) ** 4
    p[torch.rand((J, P), device="cuda", generator=g) < 0.08] = 1.0
    out = torch.empty_like(p)
    res = {"rows": J, "cols": P}
    for mode in ("pairwise", "all"):
        ops.bh_adjust(p, mode, out=out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        for i in range(3):
            ops.bh_adjust(p, mode, out=out)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(3))[1]
        res[mode] = {"ms": ms, "values_per_s": J * P / (ms * 1e-3)}
    # numpy on a few columns (the reference calls statsmodels once per column)
    cols = [0, P // 2, P - 1]
    host = p[:, cols].cpu().numpy()
    t0 = time.perf_counter()
    want = np.stack([oracle_np.bh_adjust(host[:, k]) for k in range(len(cols))], axis=1)
    dt = (time.perf_counter() - t0) / len(cols)
    ops.bh_adjust(p, "pairwise", out=out)
    got = out[:, cols].cpu().numpy()
    res["numpy_ms_per_column"] = dt * 1e3
    res["numpy_values_per_s_1core"] = J / dt
    res["bit_exact_on_sampled_columns"] = bool(np.array_equal(got.view(np.uint64), want.view(np.uint64)))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
