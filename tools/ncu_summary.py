#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: the handful of metrics the roofline uses."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:60])
    for k in KEYS:
        if k in hdr:
            print(f"  {k:72s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
    st = sorted(((float(r[i].replace(",", "") or 0), h) for i, h in enumerate(hdr)
                 if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")), reverse=True)
    for v, h in st[:7]:
        print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:32s} {v:.2f}")
