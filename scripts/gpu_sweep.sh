#!/bin/bash
# A/B sweep of run-time knobs (env) and of other BUILDS of the library (SWIFTGPU_LIB) on some workloads:
#   gpurun -- 'WLS="sphenix128 sedov128" VARIANTS="- SWIFTGPU_NO_BALANCE=1 SWIFTGPU_LIB=swift_b200/libswiftgpu_s192.so" bash scripts/gpu_sweep.sh'
mkdir -p gpurun_out
WLS=${WLS:-sedov128}
VARIANTS=${VARIANTS:--}
run() {
  local wl=$1 v=$2
  local e=""; [ "$v" != "-" ] && e=$(echo "$v" | tr ',' ' ')
  env $e timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - <<P
import json
try:
  d=json.loads(open("gpurun_out/sw.json").read().strip().splitlines()[-1])
  print("$wl [$v] ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phase_ms"].items() if v>0.05}, "cand/hit", round(d["roofline"]["candidates_per_hit"],2))
except Exception as e:
  print("$wl [$v] failed", e, open("gpurun_out/sw.err").read()[-500:])
P
}
for wl in $WLS; do for v in $VARIANTS; do run $wl $v; done; done 2>&1 | tee -a gpurun_out/sweep.log
