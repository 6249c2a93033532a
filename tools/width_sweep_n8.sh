#!/bin/bash
# BASELINE config [4] (width sweep) across the GPUs of one box: bench.py's replicas (one chr22-shaped shard per rank) at
# several region widths; one JSON line per width into gpurun_out/r2_width_n${N}.jsonl
N=${N:-8}
mkdir -p gpurun_out
: > gpurun_out/r2_width_n${N}.jsonl
for wn in ${SHAPES:-100:1000000 10000:1000000 100000:200000}; do
  w=${wn%%:*}; n=${wn##*:}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 \
    --width $w --regions $n --no-genome --no-other-ops --no-cpu-baseline 2>> gpurun_out/r2_width_n${N}.err | tail -1 >> gpurun_out/r2_width_n${N}.jsonl
done
