#!/bin/bash
# Evidence run on one B200 (gpurun): GPU tests, smoke, bench lines (own arm + reference arm), ncu launch
# lists of the same commands and one `ncu --set full` capture per headline kernel (each only after the same
# command has exited 0 without ncu).  Outputs under gpurun_out/<tag>_*; tools/profile_digest.py turns the
# .ncu-rep files into the text summaries committed under profiles/.
tag=${1:-round2}
o=gpurun_out/${tag}
python -m pytest tests -m gpu -x -q > ${o}_pytest_gpu.log 2>&1; tail -3 ${o}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > ${o}_smoke.log 2>&1; tail -1 ${o}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > ${o}_bench_ref.json 2>/dev/null
python bench.py > ${o}_bench.json 2> ${o}_bench.err; tail -c 300 ${o}_bench.json; echo
python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file ${o}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > ${o}_ncu_launches.log 2>&1
python tools/prof_kernels.py quant 400000 1000 > ${o}_prof_quant_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:quant_wide --launch-skip 2 -c 1 -o ${o}_quant -f python tools/prof_kernels.py quant 400000 1000 > ${o}_ncu_quant.log 2>&1
SD_PROF_REPS=3 python tools/prof_kernels.py fisher 200000 64 > ${o}_prof_fisher_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fisher_pairwise_binned --launch-skip 1 -c 1 -o ${o}_fisher -f python tools/prof_kernels.py fisher 200000 64 > ${o}_ncu_fisher.log 2>&1
python tools/bh_bench.py > ${o}_bh_bench.json 2>/dev/null && tools/bh_launches.sh ${o}_bh_launches.csv > ${o}_bh_launches.txt 2>&1; cat ${o}_bh_launches.txt
python tools/k1_launches.py 400000 > ${o}_k1.txt 2>&1; cat ${o}_k1.txt
cat ${o}_prof_quant_plain.log ${o}_prof_fisher_plain.log | tail -4
ls -la gpurun_out/${tag}_*.ncu-rep
