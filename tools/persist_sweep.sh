#!/bin/bash
# sweep the L2 run-ahead of the persistent engine (row pairs per warp, bytes per row)
mkdir -p gpurun_out
for cfg in "0 0" "1 4096" "1 8192" "2 4096" "2 16384" "4 4096"; do
  set -- $cfg
  CALM_B200_PF_PAIRS=$1 CALM_B200_PF_BYTES=$2 timeout 200 python bench.py --steps 32 --warmup 4 --engine 2 --no-cpu-baseline --pos0 4000 > gpurun_out/sweep.json 2>gpurun_out/sweep.err
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep.json"))
print("pairs=$1 bytes=$2 ms/tok %.3f tok/s %.1f"%(d["ms_per_step"], d["value"]), {k:(round(v["us_per_launch"],1), round(v["barrier_wait_us"],1)) for k,v in d["roofline"]["stages"].items() if k!="embed"})
PY
done
