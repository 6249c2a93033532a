#!/bin/bash
# t4 kernel variants on the bench workload: one line per setting (k_t4 ms, step value).
set -u
run() {
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); print('$*', 'k_t4_ms', round(b['by_kernel']['k_t4_ms'],4), 'k_t6_ms', round(b['by_kernel']['k_t6_ms'],4), 'value', round(b['value']/1e9,2), 'e2e', round(b['e2e']['value']/1e9,2))"
}

run VSGPU_T4_PIPE=1



run VSGPU_T4_PIPE=1 VSGPU_T4_TILE=128

run VSGPU_T4_PIPE=1
run VSGPU_T4_PIPE=1 VSGPU_T4_TILE=128
run VSGPU_T4_PIPE=1 VSGPU_T4_TILE=64
run VSGPU_T4_PIPE=0
