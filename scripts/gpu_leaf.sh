#!/bin/bash
mkdir -p gpurun_out
for TG in 5 6; do for K in tile cta; do
  SWIFTGPU_TOPGRID=$TG SWIFTGPU_HOLD=2 SWIFTGPU_LOOPS=$K timeout 300 python bench.py --workload sedov128 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/leaf_${TG}_$K.json 2> gpurun_out/leaf_${TG}_$K.err
  python - <<P
import json
try:
  d=json.loads(open("gpurun_out/leaf_${TG}_$K.json").read().strip().splitlines()[-1])
  print("tg $TG $K ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phase_ms"].items() if v>0.05}, "cand/hit", round(d["roofline"]["candidates_per_hit"],2), d["interactions_per_step"])
except Exception as e:
  print("tg $TG $K failed", e, open("gpurun_out/leaf_${TG}_$K.err").read()[-600:])
P
done; done
