#!/bin/bash
# per-kernel time and DRAM bytes of one sd_bh_adjust call (column mode, 200,000 x 2,016), via ncu
out=${1:-gpurun_out/bh_launches.csv}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"cs_|bh_|DeviceRadix" -c ${2:-9} --csv --log-file $out python tools/bh_bench.py > /dev/null 2>&1
python - "$out" <<PY
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
d = {}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki].split("(")[0][-40:]), {})[r[mi]] = float(r[vi].replace(",", ""))
tot = 0
for (i, k), v in sorted(d.items()):
    tot += v["gpu__time_duration.sum"]
    print(f"{i:3d} {k:42s} {v['gpu__time_duration.sum'] / 1e6:8.3f} ms  read {v['dram__bytes_read.sum'] / 1e9:6.2f} GB  write {v['dram__bytes_write.sum'] / 1e9:6.2f} GB")
print(f"total {tot / 1e6:.3f} ms")
PY
