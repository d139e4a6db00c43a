#!/bin/bash
# parity tests + bench for the given loop kinds (KINDS="tile cta"), prints one line each
TAG=${1:-q}
mkdir -p gpurun_out
if [ -z "$NOTEST" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
fi
for K in ${KINDS:-tile}; do for WL in ${WLS:-sedov128}; do
  SWIFTGPU_LOOPS=$K timeout 600 python bench.py --workload $WL --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_${K}_$WL.json 2> gpurun_out/${TAG}_bench_${K}_$WL.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${K}_$WL.json").read().strip().splitlines()[-1])
    print("$K $WL", "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phase_ms"].items() if v>0.05}, "cand/hit", round(d["roofline"].get("candidates_per_hit"),2), "inter", d["interactions_per_step"], "e2e", round(d["e2e"]["ms_per_step"],2), "frac", round(d["roofline"]["frac"],4))
except Exception as e:
    print("$K $WL parse failed", e); print(open("gpurun_out/${TAG}_bench_${K}_$WL.err").read()[-1500:])
P
done; done
