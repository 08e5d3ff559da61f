#!/bin/bash
# One gpurun call: GPU parity tests, the default bench line, and knob sweeps.  Outputs under gpurun_out/.
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -x -q -m gpu -k "not vs_reference_cuda" 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -8
timeout 600 python tools/sweep.py --steps 64 \
  --set base \
  --set "CALM_B200_PF=0,0,0,0,0,0" \
  --set "CALM_B200_PF=1,0,0,0,0,0" \
  --set "CALM_B200_PF=1,1,0,0,0,0" \
  --set "CALM_B200_PF=1,1,40,0,0,0" \
  --set "CALM_B200_PF=1,1,40,0,0,1" \
  --set "CALM_B200_PF=1,1,80,0,0,1" \
  --set "CALM_B200_PF=1,1,40,40,0,1" \
  --set "CALM_B200_PF=1,1,40,0,30,1" \
  --set "CALM_B200_PF=1,1,40,40,56,1" \
  --set "CALM_B200_PF=0,0,0,0,0,0;CALM_B200_EARLY=0" \
  > gpurun_out/sweep_pf.jsonl 2> gpurun_out/sweep_pf.err
cat gpurun_out/sweep_pf.jsonl; tail -3 gpurun_out/sweep_pf.err
timeout 600 python bench.py --steps 64 --warmup 8 > gpurun_out/bench_default.jsonl 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.jsonl; tail -3 gpurun_out/bench_default.err
