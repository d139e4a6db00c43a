#!/bin/bash
# ncu launch list + one --set full capture of a whole step's neighbour-loop launches of the bench command
# on <workload> (no tests, no plain bench):  gpurun -- 'bash scripts/gpu_profile.sh <tag> <workload> [skip] [count]'
TAG=${1:-r02}
WL=${2:-sphenix512}
SKIP=${3:-45}
COUNT=${4:-15}
mkdir -p gpurun_out
ARGS="--workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-resident"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_pipe|k_direct' -s $SKIP -c $COUNT \
  -f -o gpurun_out/${TAG}_full python bench.py $ARGS > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | grep $TAG
