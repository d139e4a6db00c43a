#!/bin/bash
# parity tests, then bench matrix: loops kind x reorder
TAG=${1:-ab2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
for NR in 0 1; do for K in tile cta; do
  SWIFTGPU_NO_REORDER=$NR SWIFTGPU_HOLD=2 SWIFTGPU_LOOPS=$K timeout 300 python bench.py --workload ${WL:-sedov128} --steps 5 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_${K}_$NR.json 2> gpurun_out/${TAG}_bench_${K}_$NR.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${K}_$NR.json").read().strip().splitlines()[-1])
    print("$K noreorder=$NR", "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phase_ms"].items() if v>0.05}, "cand/hit", round(d["roofline"].get("candidates_per_hit"),2), "inter", d["interactions_per_step"], d["interactions_incl_ghost_reruns"], "e2e", round(d["e2e"]["ms_per_step"],2))
except Exception as e:
    print("$K $NR parse failed", e); print(open("gpurun_out/${TAG}_bench_${K}_$NR.err").read()[-1500:])
P
done; done
