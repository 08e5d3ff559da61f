#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (oracle/_ref/libcalm_ref_cpu.so, i.e. the
unmodified reference src/infer.c compiled with the reference's flags by oracle/Makefile).

The reference ships no golden vectors or unit tests, so these fixtures are what pins
the oracle and the CUDA path on machines without /root/reference (the GPU box).  Each fixture holds,
for one seeded synthetic model (calm_b200.modelgen, seed 0) and the fixed teacher-forced token list:
  sha256     of the generated tensors (guards against a generator that drifted)
  tokens     the token list
  logits     float32 [n_keep, vocab] reference logits at the steps in `steps`
  argmax     int32 [n_tokens] reference greedy pick at every step
  margin     float32 [n_tokens] reference top-1 minus top-2
  k, v       float32 [n_layers, n_kvpos, kv_dim] KV-cache entries at positions `kvpos`

Run:  python tools/make_golden.py     (needs /root/reference; `make -C oracle all ref` is run first)
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from calm_b200 import modelgen as mg  # noqa: E402

GOLDEN_SPECS = ["tiny-fp8", "tiny-fp16", "tiny-gf4", "tiny-qwen", "tiny-llama", "tiny-moe", "tiny-moe-gf4",
                "tiny-gelu-clip", "tiny-ln", "tiny-lnpar", "tiny-mha", "tiny-bias2", "tiny-hd256", "tiny-tp8", "tiny-tp8-moe", "ring-tp"]
N_TOKENS = 24
STEPS = [0, 1, 7, 15, 23]
KVPOS = [0, 5, 23]


def model_digest(model) -> str:
    h = hashlib.sha256()
    for k in sorted(model.tensors):
        t = model.tensors[k]
        h.update(k.encode())
        h.update(t.contiguous().view(-1).view(__import__("torch").uint8).numpy().tobytes())
    return h.hexdigest()


def main():
    oracle.build(ref=True)
    ck = oracle.Checker("reference")
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    only = sys.argv[1:]  # optional: regenerate just these fixtures
    for name in GOLDEN_SPECS:
        if only and name not in only:
            continue
        spec = mg.SPECS[name]
        model = mg.HostModel(spec, seed=0)
        toks = mg.teacher_tokens(spec.vocab_size, N_TOKENS)
        logits = oracle.teacher_forced(ck, model, toks)
        srt = np.sort(logits, axis=1)
        k = np.zeros((spec.n_layers, len(KVPOS), spec.kv_dim), np.float32)
        v = np.zeros_like(k)
        for l in range(spec.n_layers):
            for i, p in enumerate(KVPOS):
                k[l, i], v[l, i] = ck.read_kv(model, l, p)
        np.savez_compressed(
            os.path.join(outdir, name + ".npz"), sha256=model_digest(model), tokens=np.array(toks, np.int32),
            steps=np.array(STEPS, np.int32), logits=logits[STEPS].astype(np.float32), argmax=logits.argmax(1).astype(np.int32),
            margin=(srt[:, -1] - srt[:, -2]).astype(np.float32), kvpos=np.array(KVPOS, np.int32), k=k, v=v)
        print(f"{name}: sigma {logits.std():.3f} min margin {(srt[:, -1] - srt[:, -2]).min():.2e}")


if __name__ == "__main__":
    main()
