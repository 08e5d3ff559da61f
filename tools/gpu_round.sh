#!/bin/bash
# One gpurun call: GPU parity tests, knob sweeps, the default bench line.  Outputs under gpurun_out/.
mkdir -p gpurun_out
set -x
timeout 1800 python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep.py --steps 64 \
  --set base \
  --set "CALM_B200_ATTN2=0" \
  > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
cat gpurun_out/sweep.jsonl; tail -3 gpurun_out/sweep.err
timeout 300 python tools/sweep.py --steps 64 --workload mistral-7b-gf4 --set base --set "CALM_B200_ATTN2=0" >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
tail -2 gpurun_out/sweep.jsonl
