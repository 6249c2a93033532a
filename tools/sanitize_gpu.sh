#!/bin/bash
# compute-sanitizer over the CUDA path (run on the GPU box): memcheck, racecheck and synccheck over __graft_entry__.smoke()
# (t2 / t3 / t4 / t5 / t6 / t7 against the oracle), one fuzz parity case, and tools/debug_t4.py at a size where every CTA of
# the persistent t4 kernel takes several tiles (all walk variants, fused and not, both coordinate widths).
out=${1:-gpurun_out/r2_sanitizer.txt}
: > "$out"
run() {
  echo "== compute-sanitizer --tool $1 ${*:2}" >> "$out"
  timeout 1200 compute-sanitizer --tool "$1" --print-limit 5 "${@:2}" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Barrier error|Invalid|Race reported|passed|failed|smoke ok|MISMATCH|ok$" | grep -v "^=========     " | cut -c1-160 | sort | uniq -c | head -12 >> "$out"
}
for tool in memcheck racecheck synccheck; do
  run $tool python __graft_entry__.py smoke
  run $tool python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_fuzz_parity_cuda and 1-True-False"
  run $tool python tools/debug_t4.py 300000
done
cat "$out"
