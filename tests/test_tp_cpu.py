"""Tensor-parallel host logic on CPU: the shape rules, and that the slices prepare_cuda takes (calm_b200/tp.py
shard_tensors restates them in numpy) really partition the model -- row shards reproduce the full q/k/v/w1/w3 rows bit for
bit, column shards of wo/w2 sum to the full projection -- checked with the oracle's own decoders for all three formats."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from calm_b200 import modelgen as mg  # noqa: E402
from calm_b200 import tp  # noqa: E402


def test_shard_dims_rules():
    d = tp.shard_dims(mg.SPECS["llama3-70b-fp8"], 8)
    assert d == {"n_heads": 8, "n_kv_heads": 1, "hidden_dim": 3584, "q_dim": 1024, "kv_dim": 128}
    assert tp.shard_dims(mg.SPECS["llama3-8b-fp8"], 2)["hidden_dim"] == 7168
    assert tp.shard_dims(mg.SPECS["llama3-8b-fp8"], 1)["n_heads"] == 32
    with pytest.raises(ValueError):
        tp.shard_dims(mg.SPECS["llama3-70b-fp8"], 16)      # 8 kv heads
    assert tp.shard_dims(mg.SPECS["mixtral-8x7b-fp8"], 8)["hidden_dim"] == 1792  # MoE: the same split inside every expert
    with pytest.raises(ValueError):
        tp.shard_dims(mg.SPECS["tiny-qwen"], 2)            # one kv head
    with pytest.raises(ValueError):
        tp.shard_dims(mg.SPECS["tiny-fp8"], 4)             # hidden 704 = 22 * 32, not a multiple of 128
    assert tp.local_spec(mg.SPECS["tiny-fp8"], 2).kv_dim == 64


def test_moe_shards_partition_every_expert(oracle_pkg):
    spec = mg.SPECS["tiny-moe"]
    model = mg.HostModel(spec, seed=0)
    shards = [tp.shard_tensors(spec, model.tensors, r, 2) for r in range(2)]
    hl = spec.hidden_dim // 2
    for leaf in ("mlp.w1", "mlp.w3"):
        full = model.tensors[f"model.layers.0.{leaf}.weight"].numpy()
        got = np.concatenate([s[f"model.layers.0.{leaf}.weight"].numpy() for s in shards], axis=1)
        assert got.shape == full.shape == (spec.n_experts, spec.hidden_dim, spec.dim) and np.array_equal(got, full)
    full = model.tensors["model.layers.0.mlp.w2.weight"].numpy()
    got = np.concatenate([s["model.layers.0.mlp.w2.weight"].numpy() for s in shards], axis=2)
    assert shards[0]["model.layers.0.mlp.w2.weight"].shape == (spec.n_experts, spec.dim, hl) and np.array_equal(got, full)


@pytest.mark.parametrize("name", ["tiny-fp8", "tiny-fp16", "tiny-gf4", "tiny-bias2"])
def test_shards_partition_the_model(oracle_pkg, name):
    spec = mg.SPECS[name]
    world = 2
    model = mg.HostModel(spec, seed=0)
    ck = oracle_pkg.Checker("port_f64")
    rng = np.random.default_rng(5)
    shards = [tp.shard_tensors(spec, model.tensors, r, world) for r in range(world)]
    d = tp.shard_dims(spec, world)
    p = "model.layers.1."
    x = rng.standard_normal(spec.dim).astype(np.float32)

    def mv(t, xin):
        a = t.numpy()
        n = a.shape[1] * (8 if spec.dbits == 4 else 1)
        return ck.matvec(spec.dbits, a, xin, n, a.shape[0])

    # row shards: concatenation of the ranks' outputs IS the full output
    for leaf, rows in (("attn.wq", d["q_dim"]), ("attn.wk", d["kv_dim"]), ("attn.wv", d["kv_dim"]), ("mlp.w1", d["hidden_dim"]), ("mlp.w3", d["hidden_dim"])):
        full = mv(model.tensors[p + leaf + ".weight"], x)
        parts = [mv(s[p + leaf + ".weight"], x) for s in shards]
        assert all(len(q) == rows for q in parts)
        assert np.array_equal(np.concatenate(parts), full), leaf
    # column shards: partials over each rank's slice of the input sum to the full projection
    for leaf, n, nl in (("attn.wo", spec.q_dim, d["q_dim"]), ("mlp.w2", spec.hidden_dim, d["hidden_dim"])):
        xin = rng.standard_normal(n).astype(np.float32)
        full = mv(model.tensors[p + leaf + ".weight"], xin)
        part = sum(mv(s[p + leaf + ".weight"], xin[r * nl:(r + 1) * nl]).astype(np.float64) for r, s in enumerate(shards))
        assert np.allclose(part, full, rtol=1e-5, atol=1e-6), leaf
    if spec.qkv_bias:
        b = model.tensors[p + "attn.wqkv.bias"].numpy()
        q = np.concatenate([s[p + "attn.wqkv.bias"].numpy()[:d["q_dim"]] for s in shards])
        k = np.concatenate([s[p + "attn.wqkv.bias"].numpy()[d["q_dim"]:d["q_dim"] + d["kv_dim"]] for s in shards])
        v = np.concatenate([s[p + "attn.wqkv.bias"].numpy()[d["q_dim"] + d["kv_dim"]:] for s in shards])
        assert np.array_equal(np.concatenate([q, k, v]), b)


def test_id_file_handoff(tmp_path):
    path = str(tmp_path / "id")
    ident = bytes(range(128))
    tp.publish_id(path, ident)
    assert tp.wait_id(path, 1.0) == ident
    with pytest.raises(TimeoutError):
        tp.wait_id(str(tmp_path / "missing"), 0.2)


@pytest.mark.parametrize("d,grid", [(4096, 256), (8192, 256), (8192, 296), (512, 32), (256, 16), (4096, 148), (5120, 160), (896, 56)])
def test_exchange_row_map_covers_every_row_once(d, grid):
    """Index arithmetic of the in-kernel all-reduce (stages.cuh tp_exchange_add / k_matres): CTA b handles row pairs
    p = b*8 + w + it*grid*8 (w = warp 0..7); the exchange addresses its partial as cell it*16 + k <-> row
    2*((it*grid + b)*8) + k, with niter = ceil((d/2 - b*8) / (grid*8)) iterations.  Every row below d must be pushed and
    summed exactly once, and both formulas must agree on which rows a CTA owns."""
    per = grid * 8
    seen = np.zeros(d, np.int32)
    for b in range(grid):
        niter = max(0, -(-(d // 2 - b * 8) // per))
        # rows the matvec loop of this CTA produces (lane 0 of warp w stores v[0], v[1] at part[it*16 + w*2 + {0,1}])
        produced = {}
        for w in range(8):
            it, p = 0, b * 8 + w
            while p < d // 2:
                produced[it * 16 + w * 2] = 2 * p
                produced[it * 16 + w * 2 + 1] = 2 * p + 1
                p += per
                it += 1
        # rows the exchange believes those cells are
        for j in range(niter * 16):
            row = 2 * (((j >> 4) * grid + b) * 8) + (j & 15)
            if row < d:
                assert produced.get(j) == row, (b, j)
                seen[row] += 1
        assert all((j >> 4) < niter for j in produced), "a produced cell lies beyond the iterations the exchange walks"
    assert np.all(seen == 1)
    assert -(-(d // 2) // per) <= 16  # TP_MAX_ITERS
