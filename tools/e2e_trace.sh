#!/bin/bash
# host-side timeline of the fused e2e call (VSGPU_TRACE) at a few chunk sizes
mkdir -p gpurun_out
: > gpurun_out/r2_e2e_trace.txt
for c in ${CHUNKS:-0 524288 262144 131072 65536}; do
  echo "== VSGPU_CHUNK_REGIONS=$c" >> gpurun_out/r2_e2e_trace.txt
  VSGPU_TRACE=1 VSGPU_CHUNK_REGIONS=$c python bench.py --steps 6 --warmup 3 --no-other-ops --no-cpu-baseline 2>gpurun_out/trace.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['e2e']['value'], d['e2e']['frac_of_pcie_ceiling'])" >> gpurun_out/r2_e2e_trace.txt
  grep "vsgpu trace. t6t4" gpurun_out/trace.err | tail -2 >> gpurun_out/r2_e2e_trace.txt
done
