#!/bin/bash
# One gpurun call: GPU parity tests, knob sweeps, the default bench line.  Outputs under gpurun_out/.
mkdir -p gpurun_out
set -x
timeout 2400 python -m pytest tests -q -m gpu -rs 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep.py --steps 64 \
  --set base \
  --set "CALM_B200_RING=0,0,0,0" \
  --set "CALM_B200_RING=0,0,0,0;CALM_B200_ATTN_CLUSTER=0" \
  --set "CALM_B200_RING=0,0,0,0;CALM_B200_ATTN2=0" \
  --set "CALM_B200_RING=4,1,3,1" \
  --set "CALM_B200_RING=3,1,3,1" \
  --set "CALM_B200_RING=3,2,3,1" \
  --set "CALM_B200_RING=2,2,2,2" \
  --set "CALM_B200_RING=2,2,4,1" \
  --set "CALM_B200_RING=4,1,4,1" \
  > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
cat gpurun_out/sweep.jsonl; tail -3 gpurun_out/sweep.err
timeout 300 python tools/sweep.py --steps 64 --workload mistral-7b-gf4 --set base --set "CALM_B200_RING=0,0,0,0" >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
tail -2 gpurun_out/sweep.jsonl
