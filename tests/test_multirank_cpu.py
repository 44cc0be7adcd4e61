"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path of bench.py — rank aggregation (max time, summed
work), the packed body-state exchange layout, and the reference arm under torchrun (rank 0 works and prints, the other
ranks exit 0 without output)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, os.environ["S2B_ROOT"])
import torch
import torch.distributed as dist
import bench
dist.init_process_group(backend="gloo")
rank = dist.get_rank()
# rank r stepped for (10 + 5 r) ms doing (1000 + r) units, e2e (0.02 + 0.01 r) s doing (500 + r) units
out = bench.aggregate_over_ranks(dist, "cpu", 10.0 + 5.0 * rank, 0.02 + 0.01 * rank, 1000 + rank, 500 + rank)
# the body-state exchange: every rank contributes its packed block, all ranks see all blocks in rank order
nb = 7
mine = torch.full(((nb + 8) * 8,), float(rank + 1), dtype=torch.float32)
allb = torch.empty(dist.get_world_size() * (nb + 8) * 8, dtype=torch.float32)
dist.all_gather_into_tensor(allb, mine)
with open(os.path.join(os.environ["S2B_OUT"], f"rank{rank}.json"), "w") as fh:  # one file per rank: stdout of two ranks interleaves
    json.dump({"rank": rank, "agg": out, "blocks": [float(allb[i * (nb + 8) * 8]) for i in range(dist.get_world_size())]}, fh)
dist.barrier()
dist.destroy_process_group()
"""


def _torchrun(args, extra_env=None, timeout=300):
    env = dict(os.environ)
    env["S2B_ROOT"] = ROOT
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    if extra_env:
        env.update(extra_env)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300)] + args
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_rank_aggregation_and_exchange_layout(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = _torchrun([str(script)], extra_env={"S2B_OUT": str(tmp_path)})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.load(open(tmp_path / f"rank{k}.json")) for k in range(2)]
    assert sorted(l["rank"] for l in lines) == [0, 1]
    for l in lines:
        total_ms_max, e2e_max, work_all, e2e_work_all = l["agg"]
        assert total_ms_max == 15.0 and abs(e2e_max - 0.03) < 1e-12
        assert work_all == 2001.0 and e2e_work_all == 1001.0
        assert l["blocks"] == [1.0, 2.0]


def test_reference_arm_under_torchrun_prints_once(reference):
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--base", "12"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "only rank 0 prints"
    line = lines[0]
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["e2e"]["h2d_bytes_per_step"] == 0
    for key in ("metric", "unit", "steps", "warmup", "ms_per_step", "higher_is_better", "config"):
        assert key in line
