#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -q -m gpu -rs > gpurun_out/pytest_gpu_full.log 2>&1; grep -v "^# CUDA" gpurun_out/pytest_gpu_full.log | tail -7
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^# CUDA" | tail -1
timeout 600 python bench.py > gpurun_out/bench_final.jsonl 2> gpurun_out/bench_final.err
tail -c 1500 gpurun_out/bench_final.jsonl; tail -3 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference_arm.jsonl 2> gpurun_out/bench_reference_arm.err; tail -c 700 gpurun_out/bench_reference_arm.jsonl
