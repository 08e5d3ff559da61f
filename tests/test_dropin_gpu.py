"""The drop-in proof on the GPU box: the UNMODIFIED reference driver (run.c, tensors.c, tokenizer.c, sampler.c) linked
against libcalm_b200.so (oracle/_ref/run_b200, built by `make -C oracle dropin`) loads a .calm file, decodes greedily on
the GPU and prints the same tokens as the reference's own CPU path (CALM_CPU=1, same binary) and as the oracle."""
import os
import re
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import ROOT  # noqa: E402

from calm_b200 import modelgen as mg  # noqa: E402

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "oracle", "_ref", "run_b200")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/run_b200 did not travel (needs /root/reference at build time)")
@pytest.mark.parametrize("name", ["tiny-llama", "tiny-gf4", "tiny-moe"])
def test_reference_driver_with_our_backend(tmp_path, oracle_pkg, name):
    spec = mg.SPECS[name]
    model = mg.HostModel(spec, seed=0)
    path = str(tmp_path / "m.calm")
    mg.write_calm(path, spec, model.tensors)
    args = [EXE, path, "-n", "12", "-t", "0", "-i", "<|t7|><|t8|>"]

    def run(env_extra):
        env = dict(os.environ, OMP_NUM_THREADS="2", **env_extra)
        r = subprocess.run(args, capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-800:]
        assert re.search(r"tok/s", r.stderr)
        return re.findall(r"<\|t(\d+)\|>", r.stdout.split("\n")[-2] if r.stdout.endswith("\n") else r.stdout)

    gpu = run({})
    cpu = run({"CALM_CPU": "1"})
    assert "# CUDA: " in subprocess.run(args, capture_output=True, text=True, timeout=300).stdout  # our prepare_cuda banner, as the reference prints
    assert len(gpu) >= 8
    assert gpu == cpu, (gpu, cpu)
    # and the oracle's greedy continuation of BOS, 7, 8
    ck = oracle_pkg.Checker("port")
    m2 = mg.HostModel(spec, seed=0)
    ck.prepare(m2)
    toks = [spec.bos_id, 7, 8]
    for i, t in enumerate(toks):
        logits = ck.forward(m2, t, i)
    gen, pos = [], len(toks)
    for _ in range(8):
        nxt = int(logits.argmax())
        gen.append(nxt)
        logits = ck.forward(m2, nxt, pos)
        pos += 1
    got = [int(t) for t in gpu]
    # the driver prints the prompt tokens after BOS, then the generated ones
    assert got[:2] == [7, 8]
    assert got[2:2 + len(gen)] == gen[:len(got) - 2]


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/run_b200 did not travel (needs /root/reference at build time)")
def test_perf_table_through_the_reference_driver(tmp_path):
    """CALM_B200_PERF=1: the unmodified reference driver, linked to our backend, ends with perf_cuda()'s per-stage table (the
    reference prints its own under a CUPTI injection, run.c:630-632, infer.cu:761-801); timings come from inside the production graph."""
    spec = mg.SPECS["tiny-llama"]
    model = mg.HostModel(spec, seed=0)
    path = str(tmp_path / "m.calm")
    mg.write_calm(path, spec, model.tensors)
    env = dict(os.environ, OMP_NUM_THREADS="2", CALM_B200_PERF="1")
    r = subprocess.run([EXE, path, "-n", "16", "-t", "0", "-i", "<|t7|>"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    assert "forward breakdown" in r.stdout and "matmul_ffn_up" in r.stdout and "GB/s" in r.stdout, r.stdout[-600:]
    rows = re.findall(r"\[(\d)\]\s+(\w+):\s+([0-9.]+)%;\s+([0-9.]+) usec/run", r.stdout)
    assert len(rows) >= 5 and abs(sum(float(x[2]) for x in rows) - 100.0) < 1.0
