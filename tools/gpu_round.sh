#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_prefill_gpu.py -x -q -s 2>&1 | tail -30 > gpurun_out/pytest_prefill.log
tail -30 gpurun_out/pytest_prefill.log
timeout 900 python tools/sweep.py --steps 64 --set base --set "CALM_B200_ATTN2=0" > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
cat gpurun_out/sweep.jsonl | cut -c1-700; tail -3 gpurun_out/sweep.err
timeout 900 python -m pytest tests -q -m gpu -x -k "not prefill" 2>&1 | tail -5
