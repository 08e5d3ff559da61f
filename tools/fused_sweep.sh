#!/bin/bash
# sweep the producer window / ring depth of the fused engine
mkdir -p gpurun_out
for cfg in "1 6" "2 6" "3 6" "6 6" "2 3"; do
  set -- $cfg
  for dbg in 0 1; do
  CALM_B200_FUSED_DBG=$dbg CALM_B200_FUSED_WINDOW=$1 CALM_B200_FUSED_SLOTS=$2 timeout 200 python bench.py --steps 32 --warmup 4 --engine 1 --no-cpu-baseline --pos0 4000 > gpurun_out/sweep.json 2>gpurun_out/sweep.err
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep.json"))
print("window=$1 slots=$2 dbg=$dbg ms/tok %.3f tok/s %.1f"%(d["ms_per_step"], d["value"]), {k:(round(v["us_per_launch"],1), round(v["barrier_wait_us"],1), round(v["load_x_us"],1), round(v["tile_wait_us"],1)) for k,v in d["roofline"]["stages"].items() if k!="embed"})
PY
  done
done
