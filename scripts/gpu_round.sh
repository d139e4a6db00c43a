#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list and one
# ncu --set full capture of the neighbour-loop kernels (density, 2 ghost re-runs, force). Outputs in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag> [workload]'
TAG=${1:-r01}
WL=${2:-sphenix128}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --workload $WL > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-resident \
  > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pipe|k_direct' -s 51 -c 17 \
  -f -o gpurun_out/${TAG}_full python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-resident \
  > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out
