#!/bin/bash
# bench alternative builds of the library: swift_b200/libswiftgpu_<tag>.so
mkdir -p gpurun_out
cp swift_b200/libswiftgpu.so /tmp/lib_default.so
run() {
  timeout 300 python bench.py --workload sedov128 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/lib_$1.json 2> gpurun_out/lib_$1.err
  python - <<P
import json
try:
  d=json.loads(open("gpurun_out/lib_$1.json").read().strip().splitlines()[-1])
  print("$1 ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["phase_ms"].items() if v>0.05})
except Exception as e:
  print("$1 failed", e, open("gpurun_out/lib_$1.err").read()[-800:])
P
}
run default
for f in swift_b200/libswiftgpu_*.so; do
  t=$(basename $f .so); t=${t#libswiftgpu_}
  [ "$t" = "host" ] && continue
  cp $f swift_b200/libswiftgpu.so; run $t
done
cp /tmp/lib_default.so swift_b200/libswiftgpu.so
