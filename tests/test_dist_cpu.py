"""CPU tests of the multi-process plumbing (no GPU): bench.py's reference arm under a 2-rank launch, and the
rank aggregation bench.py uses (barrier + max-over-ranks of the step time, sum of the per-rank work) over gloo."""
import json
import os
import socket
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import ROOT  # noqa: E402


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_reference_arm_two_ranks(oracle_pkg):
    """torchrun-style launch with WORLD_SIZE=2: rank 0 runs the CPU reference and prints ONE JSON line, rank 1 exits 0
    without work (bench.py contract for --impl reference)."""
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "3",
                                       "--workload", "tiny-llama"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-500:] for o in outs]
    assert outs[1][0].strip() == ""
    lines = [l for l in outs[0][0].splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tok/s" and d["value"] > 0 and d["steps"] == 3
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
import bench
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
agg = bench.RankAggregator(dist, device="cpu")
agg.barrier()
ms = agg.max_over_ranks(10.0 + 5.0 * rank)          # slowest rank decides
total = agg.whole_job_rate(steps=8, ms=ms)           # every rank did 8 steps
if rank == 0:
    print(ms, total)
dist.destroy_process_group()
"""


def test_rank_aggregation_gloo(tmp_path):
    port = free_port()
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1][-800:] for o in outs]
    ms, total = [float(x) for x in outs[0][0].split()]
    assert ms == 15.0                       # max over ranks
    assert abs(total - 2 * 8 / 0.015) < 1e-6  # whole-job tokens/s: all ranks' steps over the slowest rank's time
