#!/usr/bin/env python
"""Print the SASS of an `ncu --page source --csv --print-source sass` dump with executed
counts per unit of work and stall samples.  usage: sass_profile.py dump.csv units [min_count]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])
hdr = rows[1]
ie, src, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = 0
for r in rows[2:]:
    if len(r) <= ie or not r[ie].isdigit():
        if r and r[0] == "Kernel Name":
            break
        continue
    n = int(r[ie]); tot += n
    print(f"{n / units:8.2f} {int(r[smp]):6d}  {r[src]}")
print("total warp instructions", tot, "per unit", tot / units)
