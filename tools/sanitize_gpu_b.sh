#!/bin/bash
# compute-sanitizer over what round 2 added after tools/sanitize_gpu.sh ran: the spilling instance of k_t4p (ample and too little
# scratch), the t5 kernels that keep hit codes, and the t5 row rendering.
out=${1:-gpurun_out/r2_sanitizer_b.txt}
: > "$out"
run() {
  echo "== compute-sanitizer --tool $1 ${*:2}" >> "$out"
  timeout 900 compute-sanitizer --tool "$1" --print-limit 5 "${@:2}" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Barrier error|Invalid|Race reported|passed|failed|smoke ok|MISMATCH|ok$" | grep -v "^=========     " | cut -c1-160 | sort | uniq -c | head -12 >> "$out"
}
for tool in memcheck racecheck synccheck; do
  run $tool python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_fuzz_parity_cuda and 1-True-False and spill"
  run $tool python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_t5_rows_rendered_on_device_cuda and False-True"
  run $tool python __graft_entry__.py smoke
done
cat "$out"
