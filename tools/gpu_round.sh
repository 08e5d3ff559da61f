#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -q -m gpu -x -k "golden or long_context or fp8_kv or rolling or kv_only or greedy or full_size" 2>&1 | grep -v "^# CUDA" | tail -5
timeout 300 python tools/sweep.py --steps 64 --set base --set base > gpurun_out/sweep_attn_batched_fold.jsonl 2> gpurun_out/sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/sweep_attn_batched_fold.jsonl'):
    r=json.loads(l); print(r['cfg'], r['ms_per_token'], r['us_per_launch'], r['attn_dbg_ns'])
PY
tail -3 gpurun_out/sweep.err
