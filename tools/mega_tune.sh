#!/bin/bash
# A/B the pipeline kernel's knobs on the GPU box: bash tools/mega_tune.sh "<env assignments>" ...
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --steps 300 --warmup 10 --no-cpu --no-sweep > gpurun_out/mt_$i.json 2> gpurun_out/mt_$i.err
  python - "$cfg" gpurun_out/mt_$i.json <<PY
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print("%-60s value %9.0f  us/step %6.2f  two_pass %9.0f  k_us %6.2f"%(sys.argv[1],d["value"],d["ms_per_step"]*1e3,d["two_pass"]["value"],d["roofline"]["kernel_ms"]*1e3))
except Exception as e:
    print(sys.argv[1],"FAILED",e)
PY
done
