#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -q -m gpu -x -s -k "ring_fed or reference_widths or prefill_at_reference" 2>&1 | grep -v "^# CUDA" | tail -14
timeout 300 python tools/sweep.py --steps 64 --workload mistral-7b-gf4 --set base --set "CALM_B200_RING=2,2,0,0" --set "CALM_B200_RING=0,0,0,0" > gpurun_out/sweep_gf4.jsonl 2> gpurun_out/sweep_gf4.err
cat gpurun_out/sweep_gf4.jsonl | cut -c1-420; tail -3 gpurun_out/sweep_gf4.err
timeout 300 python tools/sweep.py --steps 64 --set base --set "CALM_B200_ATTN_SPLIT_MUL=2" > gpurun_out/sweep_attn.jsonl 2> gpurun_out/sweep_attn.err
cat gpurun_out/sweep_attn.jsonl | cut -c1-700; tail -3 gpurun_out/sweep_attn.err
