#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -q -m gpu -x -s -k "ring_fed or reference_widths or prefill_at_reference" 2>&1 | grep -v "^# CUDA" | tail -12
timeout 300 python tools/sweep.py --steps 64 --workload mistral-7b-gf4 --set base --set "CALM_B200_GF4_MMA=0" > gpurun_out/sweep_gf4.jsonl 2> gpurun_out/sweep_gf4.err
cat gpurun_out/sweep_gf4.jsonl | cut -c1-420; tail -3 gpurun_out/sweep_gf4.err
timeout 600 python bench.py --steps 64 --warmup 8 > gpurun_out/bench_default.jsonl 2> gpurun_out/bench_default.err
tail -c 2600 gpurun_out/bench_default.jsonl; tail -3 gpurun_out/bench_default.err
# launch list (cold, serialised: shares only) and one full capture of the stage kernels of a middle layer
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -s 400 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ffn_up_ring|k_matres_ring|k_attn2|k_qkv" -s 80 -c 5 -o gpurun_out/prof_layer python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
