#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep into the tracked text summaries under profiles/.

    python tools/profile_digest.py gpurun_out/r1c_quant.ncu-rep profiles/r1_quant_wide  3.125e6 "warp-rows (256 columns)"
writes <out>.summary.txt (key metrics + stall reasons) and <out>.sass.txt (executed warp
instructions per unit of work and stall samples per SASS line)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rep, out, units, unit_name = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
    raw = out + ".raw.csv.tmp"
    src = out + ".src.csv.tmp"
    subprocess.run(f"ncu -i {rep} --page raw --csv > {raw} 2>/dev/null", shell=True, check=True)
    subprocess.run(f"ncu -i {rep} --page source --csv --print-source sass > {src} 2>/dev/null", shell=True, check=True)
    with open(out + ".summary.txt", "w") as f:
        f.write(f"# digest of {os.path.basename(rep)} (ncu --set full --clock-control none)\n")
        f.write(subprocess.run([sys.executable, os.path.join(HERE, "ncu_summary.py"), raw], capture_output=True,
                               text=True).stdout)
    with open(out + ".sass.txt", "w") as f:
        f.write(f"# executed warp instructions per unit ({unit_name}) | stall samples | SASS\n")
        f.write(subprocess.run([sys.executable, os.path.join(HERE, "sass_profile.py"), src, units],
                               capture_output=True, text=True).stdout)
    os.remove(raw)
    os.remove(src)


if __name__ == "__main__":
    main()
