#!/bin/bash
# bench the tile loops with different hold depths
mkdir -p gpurun_out
for HOLD in 3 2 1; do
  SWIFTGPU_HOLD=$HOLD timeout 300 python bench.py --workload sedov128 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/hold_$HOLD.json 2> gpurun_out/hold_$HOLD.err
  python - <<P
import json
d=json.loads(open("gpurun_out/hold_$HOLD.json").read().strip().splitlines()[-1])
print("hold $HOLD ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phase_ms"].items() if v>0.05})
P
done
