#!/bin/bash
# device-resident bench line under the t4 kernel variants (run on the GPU box)
run() {
  echo "== $*"
  env "$@" python bench.py --steps 20 --warmup 3 --no-other-ops --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); b=d['by_kernel']; print(json.dumps({'ms_per_step':round(d['ms_per_step'],4),'k_t6_ms':round(b['k_t6_ms'],4),'k_t4_ms':round(b['k_t4_ms'],4),'e2e':round(d['e2e']['value']/1e9,3)}))"
}
run VSGPU_T4_STAGED=0 VSGPU_T4_ROW64=0
run VSGPU_T4_STAGED=0
run VSGPU_T4_STAGED=0 VSGPU_BUCKET_TARGET=4
run VSGPU_T4_STAGED=0 VSGPU_BUCKET_TARGET=2
run VSGPU_T4S_TILE=128
run VSGPU_T4S_TILE=128 VSGPU_BUCKET_TARGET=4
run VSGPU_T4S_TILE=128 VSGPU_BUCKET_TARGET=2
run VSGPU_T4S_TILE=64 VSGPU_BUCKET_TARGET=4
run VSGPU_T4S_TILE=256 VSGPU_BUCKET_TARGET=4
