#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python tools/sweep.py --steps 64 --set base --set "CALM_B200_RING_QKV=1" --set "CALM_B200_RING_QKV=2" > gpurun_out/sweep_qkv.jsonl 2> gpurun_out/sweep_qkv.err
cat gpurun_out/sweep_qkv.jsonl | cut -c1-700; tail -3 gpurun_out/sweep_qkv.err
CALM_B200_RING_QKV=1 timeout 300 python -m pytest tests -q -m gpu -x -k "golden or ring_fed or full_size" 2>&1 | tail -3
