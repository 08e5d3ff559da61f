#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -q -m gpu -rs > gpurun_out/pytest_gpu_full.log 2>&1; grep -v "^# CUDA" gpurun_out/pytest_gpu_full.log | tail -12
timeout 600 python bench.py > gpurun_out/bench_final.jsonl 2> gpurun_out/bench_final.err
tail -c 1800 gpurun_out/bench_final.jsonl; tail -3 gpurun_out/bench_final.err
for wl in mistral-7b-gf4 mixtral-8x7b-fp8 llama3-8b-fp16; do
  timeout 400 python bench.py --workload $wl --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench_$wl.jsonl 2> gpurun_out/bench_$wl.err; tail -c 900 gpurun_out/bench_$wl.jsonl; tail -2 gpurun_out/bench_$wl.err
done
timeout 400 python bench.py --kvbits 8 --steps 64 --warmup 8 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_llama3-8b-fp8_kv8.jsonl 2> gpurun_out/bench_kv8.err; tail -c 600 gpurun_out/bench_llama3-8b-fp8_kv8.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pf_" --csv --log-file gpurun_out/launches_prefill.csv python tools/prefill_bench.py --layers 2 --reps 1 > gpurun_out/prefill_ncu.log 2>&1; tail -2 gpurun_out/prefill_ncu.log
timeout 300 python tools/prefill_bench.py --layers 8 > gpurun_out/prefill_8l.json 2>&1; tail -1 gpurun_out/prefill_8l.json
timeout 600 python bench.py --workload llama3-70b-fp8 --steps 32 --warmup 4 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_llama3-70b-fp8.jsonl 2> gpurun_out/bench_70b.err; tail -c 700 gpurun_out/bench_llama3-70b-fp8.jsonl; tail -2 gpurun_out/bench_70b.err
