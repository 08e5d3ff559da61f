#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -q -m gpu -x -s -k "ring_fed or reference_widths or prefill_at_reference or dropin" 2>&1 | grep -v "^# CUDA" | tail -14
timeout 300 python tools/sweep.py --steps 64 --workload mistral-7b-gf4 --set base --set "CALM_B200_MMA=0" --set "CALM_B200_MMA_RES=0" > gpurun_out/sweep_gf4.jsonl 2> gpurun_out/sweep_gf4.err
cat gpurun_out/sweep_gf4.jsonl | cut -c1-420; tail -3 gpurun_out/sweep_gf4.err
timeout 300 python tools/sweep.py --steps 64 --set base --set "CALM_B200_MMA_RES=1" --set "CALM_B200_MMA=1;CALM_B200_MMA_RES=1" > gpurun_out/sweep_fp8_mma.jsonl 2> gpurun_out/sweep_fp8.err
cat gpurun_out/sweep_fp8_mma.jsonl | cut -c1-420; tail -3 gpurun_out/sweep_fp8.err
timeout 300 python tools/sweep.py --steps 64 --workload llama3-8b-fp16 --set base --set "CALM_B200_MMA=0" --set "CALM_B200_MMA_RES=1" > gpurun_out/sweep_fp16.jsonl 2> gpurun_out/sweep_fp16.err
cat gpurun_out/sweep_fp16.jsonl | cut -c1-420; tail -3 gpurun_out/sweep_fp16.err
