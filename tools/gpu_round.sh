#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -q -m gpu -rs > gpurun_out/pytest_gpu_full.log 2>&1; grep -v "^# CUDA" gpurun_out/pytest_gpu_full.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^# CUDA" | tail -2
timeout 600 python bench.py --workload llama3-70b-fp8 --steps 32 --warmup 4 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_llama3-70b-fp8.jsonl 2> gpurun_out/bench_70b.err; tail -c 900 gpurun_out/bench_llama3-70b-fp8.jsonl; tail -2 gpurun_out/bench_70b.err
