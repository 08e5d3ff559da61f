#!/bin/bash
# Decompose the fused kernel's token time: full run, stream-only (consumers skip the math), math-only (producer skips the copies).
# usage: tools/fused_experiments.sh "0 1 2"
mkdir -p gpurun_out
python -c "from calm_b200 import lib; L=lib.load(); print('grid barrier: %.2f us' % L.calm_b200_barrier_bench(2000))"
for dbg in ${1:-0 1 2}; do
  for pos in 4000; do
    CALM_B200_FUSED_DBG=$dbg timeout 200 python bench.py --steps 48 --warmup 6 --engine ${ENGINE:-2} --no-cpu-baseline --pos0 $pos > gpurun_out/exp_d${dbg}_p${pos}.json 2>gpurun_out/exp.err
    python - <<PY
import json
d=json.load(open("gpurun_out/exp_d${dbg}_p${pos}.json"))
print("dbg=$dbg pos=$pos  ms/tok %.3f  tok/s %.1f  e2e %.1f stages(us total, barrier wait, load_x, tile wait):"%(d["ms_per_step"], d["value"], d["e2e"]["value"]), {k:(round(v["us_per_launch"],1), round(v["barrier_wait_us"],1), round(v["load_x_us"],1), round(v["tile_wait_us"],1)) for k,v in d["roofline"]["stages"].items()})
PY
  done
done
