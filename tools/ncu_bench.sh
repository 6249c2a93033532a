#!/bin/bash
# ncu evidence for the bench's kernels (run on the GPU box, one GPU): the launch list of a short bench run
# (gpu__time_duration per launch) and one --set full capture of the t4 kernel in both forms (fused with t6, alone).
tag=${1:-r2}
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-other-ops --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_t4p -s 8 -c 2 -o gpurun_out/${tag}_t4p \
    python bench.py --steps 2 --warmup 3 --no-other-ops --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
ncu --set full --clock-control none -k regex:k_t6 -s 4 -c 1 -o gpurun_out/${tag}_t6 \
    python bench.py --steps 2 --warmup 3 --no-other-ops --no-cpu-baseline > gpurun_out/${tag}_ncu6.log 2>&1
ls -la gpurun_out/ | grep ${tag}
