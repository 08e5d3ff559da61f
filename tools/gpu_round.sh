#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python tools/sweep.py --steps 64 --set base --set "CALM_B200_RING_HINT=0" --set "CALM_B200_EARLY=0" --set base > gpurun_out/sweep_hint.jsonl 2> gpurun_out/sweep.err
cat gpurun_out/sweep_hint.jsonl | cut -c1-330; tail -3 gpurun_out/sweep.err
