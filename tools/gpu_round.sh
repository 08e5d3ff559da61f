#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python tools/sweep.py --steps 64 --set base --set "CALM_B200_RING_RES_U=2;CALM_B200_RING=2,2,3,8" --set "CALM_B200_RING_RES_U=2;CALM_B200_RING=2,2,4,16" --set "CALM_B200_RING_RES_U=2;CALM_B200_RING=2,2,2,8" --set "CALM_B200_RING=2,2,2,8" --set base > gpurun_out/sweep_w2.jsonl 2> gpurun_out/sweep.err
cat gpurun_out/sweep_w2.jsonl | cut -c1-330; tail -3 gpurun_out/sweep.err
