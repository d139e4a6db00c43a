#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for SP in 0 28; do
  SWIFTGPU_SPARSE=$SP timeout 600 python bench.py --workload sphenix128a5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/act_$SP.json 2> gpurun_out/act_$SP.err
  python - <<P
import json
try:
  d=json.loads(open("gpurun_out/act_$SP.json").read().strip().splitlines()[-1])
  print("sparse_thr $SP ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phase_ms"].items() if v>0.02}, d["interactions_per_step"], "cand/hit", round(d["roofline"]["candidates_per_hit"],2))
except Exception as e:
  print("$SP failed", e, open("gpurun_out/act_$SP.err").read()[-1500:])
P
done
