#!/bin/bash
# A/B harness for the large-batch path: bench.py --pairs N against every library under rgbd_gan_b200/lib/variants/
# usage (on the GPU box): bash tools/ab.sh [pairs] [extra bench args]
P=${1:-256}; shift
mkdir -p gpurun_out
for so in rgbd_gan_b200/lib/variants/*.so; do
  name=$(basename $so .so)
  for rep in 1 2; do
  RGBD_B200_LIB=$PWD/$so python bench.py --pairs $P --steps 200 --warmup 10 --no-cpu --no-sweep "$@" > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<PY
import json,sys
try:
    d=json.load(open("gpurun_out/ab_%s.json"%sys.argv[1]))
    print("%-16s value %9.0f  us/step %7.2f  kernel_us %7.2f  frac %.3f"%(sys.argv[1],d["value"],d["ms_per_step"]*1e3,d["roofline"]["kernel_ms"]*1e3,d["roofline"]["step"]["frac"]))
except Exception as e:
    print(sys.argv[1],"FAILED",e)
PY
  done
done
