#!/bin/bash
# Round-end evidence run on one B200 (gpurun): GPU tests, bench lines (own arm + reference arm),
# the ncu launch lists of the same commands.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r2}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>/dev/null
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.json
python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/bh_bench.py > gpurun_out/${tag}_bh_bench.json 2>/dev/null && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_bh_launches.csv python tools/bh_bench.py > /dev/null 2>&1
python tools/prof_kernels.py misc > gpurun_out/${tag}_misc.log 2>&1; tail -8 gpurun_out/${tag}_misc.log
