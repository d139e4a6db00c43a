#!/bin/bash
# sweep of run-time knobs of the tile loops on one workload
mkdir -p gpurun_out
WL=${WL:-sedov128}
run() {
  env $1 timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - <<P
import json
try:
  d=json.loads(open("gpurun_out/sw.json").read().strip().splitlines()[-1])
  print("$1 ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phase_ms"].items() if v>0.05})
except Exception as e:
  print("$1 failed", e, open("gpurun_out/sw.err").read()[-500:])
P
}
for H in 1 2 3; do run SWIFTGPU_HOLD=$H; done
for S in 0 20 28 40 64; do run SWIFTGPU_SPARSE=$S; done
