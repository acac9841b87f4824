#!/bin/bash
# A/B harness: run bench.py against every tuning variant under rgbd_gan_b200/lib/variants/
# usage (on the GPU box): bash tools/tune.sh [extra bench args]
mkdir -p gpurun_out
for so in rgbd_gan_b200/lib/variants/*.so; do
  name=$(basename $so .so)
  RGBD_B200_LIB=$PWD/$so python bench.py --steps 200 --warmup 10 --no-cpu --no-sweep "$@" > gpurun_out/tune_$name.json 2> gpurun_out/tune_$name.err
  python - "$name" <<PY
import json,sys
try:
    d=json.load(open("gpurun_out/tune_%s.json"%sys.argv[1]))
    print("%-8s value %9.0f  us/step %6.2f  two_pass %9.0f  k2_us %6.2f fwdk2_us %6.2f share %.2f"%(sys.argv[1],d["value"],d["ms_per_step"]*1e3,d["two_pass"]["value"],d["roofline"]["kernel_ms"]*1e3,d["roofline"]["kernel_ms_loss_only_variant"]*1e3,d["roofline"]["kernel_share_of_step"]))
except Exception as e:
    print(sys.argv[1],"FAILED",e)
PY
done
