#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_final3.jsonl 2> gpurun_out/bench_final3.err; cut -c1-300 gpurun_out/bench_final3.jsonl; tail -2 gpurun_out/bench_final3.err
python - <<'PY'
import json
r=json.loads(open('gpurun_out/bench_final3.jsonl').read().strip().splitlines()[-1])
print(r['value'], r['e2e'], r['prefill'], r['ref_cuda'], r['frac_of_peak_whole_token'])
PY
