#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python tools/sweep.py --steps 64 \
  --set base \
  --set "CALM_B200_RING=2,2,0,0" \
  --set "CALM_B200_RING=2,2,2,8" \
  --set "CALM_B200_RING=2,2,3,16" \
  --set "CALM_B200_RING=0,0,2,16" \
  --set "CALM_B200_ATTN2=0" \
  > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
cat gpurun_out/sweep.jsonl; tail -3 gpurun_out/sweep.err
timeout 600 python -m pytest tests -q -m gpu -x -k "golden or scale or parity" 2>&1 | tail -5
