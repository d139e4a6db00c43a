#!/bin/bash
# A/B on one GPU box: parity tests with the default loops, then bench with tile and cta loops.
TAG=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
for K in ${KINDS:-tile cta}; do
  SWIFTGPU_LOOPS=$K timeout 300 python bench.py --workload sedov128 --steps 5 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_$K.json 2> gpurun_out/${TAG}_bench_$K.err
  echo "bench $K exit $?"
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$K.json").read().strip().splitlines()[-1])
    print("$K", "ms/step", round(d["ms_per_step"],3), d["phase_ms"], "cand/hit", d["roofline"].get("candidates_per_hit"), "inter", d["interactions_per_step"], d["interactions_incl_ghost_reruns"])
except Exception as e:
    print("$K parse failed", e); print(open("gpurun_out/${TAG}_bench_$K.err").read()[-2000:])
P
done
