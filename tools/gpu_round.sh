#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -q -m gpu -x -k "golden or long_context or fp8_kv or full_size or rolling or kv_only or greedy" 2>&1 | grep -v "^# CUDA" | tail -3
timeout 600 python tools/sweep.py --steps 64 --set base --set base > gpurun_out/sweep_attn_fast.jsonl 2> gpurun_out/sweep.err
cat gpurun_out/sweep_attn_fast.jsonl | cut -c1-700; tail -3 gpurun_out/sweep.err
