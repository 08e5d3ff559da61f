#!/usr/bin/env python
"""Debug helper: run a few teacher-forced tokens of one spec on both CUDA engines and print the differences.
   python tools/debug_fused.py tiny-qwen [ntokens] [--greedy]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from calm_b200 import lib  # noqa: E402
from calm_b200 import modelgen as mg  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "tiny-qwen"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
spec = mg.SPECS[name]
host = mg.HostModel(spec, seed=0)
toks = mg.teacher_tokens(spec.vocab_size, n)
out = {}
for eng in (0, 1):
    print("engine", eng, flush=True)
    with lib.DeviceModel(spec, host.tensors, engine=eng) as dm:
        res = []
        for i, t in enumerate(toks):
            res.append(dm.forward(t, i))
            print("  token", i, "ok", flush=True)
        out[eng] = np.stack(res)
        if "--greedy" in sys.argv:
            print("  greedy", list(dm.decode_greedy(toks[0], 0, n)), flush=True)
print("max |e0-e1|", np.abs(out[0] - out[1]).max(), "sigma", out[0].std())
