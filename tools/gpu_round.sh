#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -q -m gpu -x -k "golden or ring_fed or reference_widths or gf4 or moe" 2>&1 | grep -v "^# CUDA" | tail -5
timeout 300 python tools/sweep.py --steps 64 --set base > gpurun_out/sweep_gf4_group_sums.jsonl 2> gpurun_out/sweep.err
timeout 300 python tools/sweep.py --steps 64 --workload mistral-7b-gf4 --set base >> gpurun_out/sweep_gf4_group_sums.jsonl 2>> gpurun_out/sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/sweep_gf4_group_sums.jsonl'):
    r=json.loads(l); print(r['cfg'], r['ms_per_token'], r['us_per_launch'])
PY
tail -3 gpurun_out/sweep.err
