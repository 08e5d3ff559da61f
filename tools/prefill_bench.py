#!/usr/bin/env python
"""Time forward_prefill_cuda on a workload shape (CUDA events on the library's stream); meant to be run under
`ncu --metrics gpu__time_duration.sum -k regex:k_pf_` for the per-kernel split as well.
  python tools/prefill_bench.py [--workload llama3-8b-fp8] [--layers 4] [--tokens 2048]"""
import argparse
import json
import os
import sys
from dataclasses import replace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from calm_b200 import lib  # noqa: E402
from calm_b200 import modelgen as mg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="llama3-8b-fp8")
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--tokens", type=int, default=2048)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
os.environ.setdefault("CALM_B200_QUIET", "1")
spec = replace(mg.SPECS[a.workload], n_layers=a.layers)
L = lib.load()
tensors = mg.generate(spec, 0, device="cuda")
torch.cuda.synchronize()
toks = np.array(mg.teacher_tokens(spec.vocab_size, a.tokens), np.int32)
with lib.DeviceModel(spec, tensors, seq_len=4096) as dm:
    served = dm.prefill(toks[:256], 0)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(a.reps):
        L.calm_b200_timer_start()
        dm.prefill(toks, 0)
        best = min(best, L.calm_b200_timer_stop())
mm = spec.n_layers * ((spec.q_dim + 2 * spec.kv_dim) * spec.dim + spec.dim * spec.q_dim + 3 * spec.hidden_dim * spec.dim)
flops = 2.0 * a.tokens * mm + 4.0 * spec.n_layers * spec.q_dim * a.tokens * (a.tokens + 1) / 2
print(json.dumps({"workload": spec.name, "layers": a.layers, "tokens": a.tokens, "served": served, "ms": best, "ms_per_layer": best / a.layers,
                  "tok_s": a.tokens / best * 1e3, "tflops": flops / best / 1e9}))
