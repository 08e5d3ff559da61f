"""Host side of tensor parallelism (the reference is single-device, src/infer.cu:79).

The split itself happens inside prepare_cuda (csrc/engine.cu tp_shard_model): rank r of N keeps query heads
[r*n_heads/N, (r+1)*n_heads/N), the kv heads they attend to, and FFN rows [r*hidden/N, (r+1)*hidden/N); the output
projections wo / w2 are split by COLUMN, so each rank produces a partial dim-vector that one all-reduce sums.
This module holds what a host needs around that: the shape rules (shard_dims, the same checks the library aborts
on), numpy restatements of the slices (shard_tensors; used by the CPU tests to prove sum-of-partials == full
matvec on the oracle's decoders), the NCCL id hand-off by file, and a small worker (python -m calm_b200.tp) that
the 2-GPU parity test launches once per rank.
"""
import argparse
import os
import sys
import time
from dataclasses import replace
from typing import Dict

import numpy as np

from . import modelgen as mg


def shard_dims(spec: mg.ModelSpec, world: int) -> Dict[str, int]:
    """Per-rank shapes, or ValueError with the reason prepare_cuda would abort for."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if spec.n_heads % world or spec.n_kv_heads % world:
        raise ValueError(f"{world} ranks do not divide heads {spec.n_heads}/{spec.n_kv_heads}")
    if spec.hidden_dim % (32 * world):
        raise ValueError(f"hidden_dim {spec.hidden_dim} is not a multiple of 32*{world}")
    d = {"n_heads": spec.n_heads // world, "n_kv_heads": spec.n_kv_heads // world, "hidden_dim": spec.hidden_dim // world}
    d["q_dim"], d["kv_dim"] = d["n_heads"] * spec.head_dim, d["n_kv_heads"] * spec.head_dim
    if d["q_dim"] % 32 or d["kv_dim"] % 32:
        raise ValueError("per-rank q_dim and kv_dim must be multiples of 32")
    return d


def local_spec(spec: mg.ModelSpec, world: int) -> mg.ModelSpec:
    d = shard_dims(spec, world)
    return replace(spec, name=f"{spec.name}/tp{world}", n_heads=d["n_heads"], n_kv_heads=d["n_kv_heads"], hidden_dim=d["hidden_dim"])


def _cols(t, c0: int, n: int, dbits: int):
    """Column range [c0, c0+n) of a row-major quantised matrix (last axis); gf4 packs 8 weights per stored u32."""
    per = 8 if dbits == 4 else 1
    assert c0 % per == 0 and n % per == 0
    return t[..., c0 // per:(c0 + n) // per].contiguous()


def shard_tensors(spec: mg.ModelSpec, tensors: Dict[str, "object"], rank: int, world: int) -> Dict[str, "object"]:
    """The tensors rank `rank` computes with, as a model of the LOCAL spec (same slices as tp_shard_model)."""
    d = shard_dims(spec, world)
    ql, kl, hl = d["q_dim"], d["kv_dim"], d["hidden_dim"]
    q_dim, kv_dim = spec.q_dim, spec.kv_dim
    out = {}
    for name, t in tensors.items():
        if name.endswith("attn.wq.weight"):
            t = t[rank * ql:(rank + 1) * ql].contiguous()
        elif name.endswith("attn.wk.weight") or name.endswith("attn.wv.weight"):
            t = t[rank * kl:(rank + 1) * kl].contiguous()
        elif name.endswith("mlp.w1.weight") or name.endswith("mlp.w3.weight"):
            t = t[..., rank * hl:(rank + 1) * hl, :].contiguous()  # MoE: the same row range of every expert
        elif name.endswith("attn.wo.weight"):
            t = _cols(t, rank * ql, ql, spec.dbits)
        elif name.endswith("mlp.w2.weight"):
            t = _cols(t, rank * hl, hl, spec.dbits)
        elif name.endswith("attn.wqkv.bias"):
            import torch

            t = torch.cat([t[rank * ql:(rank + 1) * ql], t[q_dim + rank * kl:q_dim + (rank + 1) * kl],
                           t[q_dim + kv_dim + rank * kl:q_dim + kv_dim + (rank + 1) * kl]]).contiguous()
        out[name] = t
    return out


# ------------------------------------------------------------------------------------------------
# NCCL id hand-off without torch.distributed (the drop-in host is a C program; a file is the lowest common denominator)

def publish_id(path: str, ident: bytes) -> None:
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(ident)
    os.replace(tmp, path)  # atomic: readers see nothing or all 128 bytes


def wait_id(path: str, timeout_s: float = 120.0) -> bytes:
    t0 = time.time()
    while time.time() - t0 < timeout_s:
        if os.path.exists(path) and os.path.getsize(path) == 128:
            with open(path, "rb") as f:
                return f.read()
        time.sleep(0.05)
    raise TimeoutError(f"no NCCL id at {path} after {timeout_s} s")


def main(argv=None):
    """One tensor-parallel rank: teacher-forced forward over the fixture token list, logits saved by rank 0."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--spec", required=True)
    ap.add_argument("--rank", type=int, required=True)
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--idfile", required=True)
    ap.add_argument("--tokens", type=int, default=24)
    ap.add_argument("--out", required=True)
    ap.add_argument("--greedy", type=int, default=0, help="also run the device-resident greedy loop for this many tokens")
    ap.add_argument("--seq-len", type=int, default=None)
    a = ap.parse_args(argv)

    import faulthandler

    faulthandler.enable()
    os.environ.setdefault("CALM_B200_QUIET", "1")
    from . import lib

    spec = mg.SPECS[a.spec]
    L = lib.load()
    L.calm_b200_set_device(a.rank)
    if a.rank == 0:
        ident = lib.tp_unique_id()
        publish_id(a.idfile, ident)
    else:
        ident = wait_id(a.idfile)
    model = mg.HostModel(spec, seed=0)
    dm = lib.DeviceModel(spec, model.tensors, device=a.rank, tp=(a.rank, a.world, ident), seq_len=a.seq_len)
    toks = mg.teacher_tokens(spec.vocab_size, a.tokens)
    logits = np.stack([dm.forward(int(t), i) for i, t in enumerate(toks)])
    greedy = dm.decode_greedy(int(toks[0]), 0, a.greedy) if a.greedy else np.zeros(0, np.int32)
    np.savez(f"{a.out}.rank{a.rank}.npz", logits=logits, greedy=greedy, world=L.calm_b200_tp_world(), mode=L.calm_b200_tp_mode(), launches=L.calm_b200_launch_count())
    dm.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
