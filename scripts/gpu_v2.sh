#!/bin/bash
# experiment: density compiled for 2 CTAs/SM with a 6-stage ring (libswiftgpu_v2.so) vs default
mkdir -p gpurun_out
cp swift_b200/libswiftgpu.so /tmp/lib_default.so
run() {
  SWIFTGPU_HOLD=$2 timeout 300 python bench.py --workload sedov128 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/v2_$1_$2.json 2> gpurun_out/v2_$1_$2.err
  python - <<P
import json
try:
  d=json.loads(open("gpurun_out/v2_$1_$2.json").read().strip().splitlines()[-1])
  print("$1 hold $2 ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phase_ms"].items() if v>0.05})
except Exception as e:
  print("$1 $2 failed", e, open("gpurun_out/v2_$1_$2.err").read()[-800:])
P
}
run default 2
cp swift_b200/libswiftgpu_v2.so swift_b200/libswiftgpu.so
for H in 2 3 4 5; do run v2 $H; done
cp /tmp/lib_default.so swift_b200/libswiftgpu.so
