"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the sharded single-instance solve (column
blocks of A, in-kernel peer-memory all-reduce) against the single-GPU engine, launched through torchrun."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("case", ["small", "mcf"])
def test_two_gpu_solve_matches_single_gpu(case):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", "dist_check.py"), case]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert lines, out.stdout[-2000:] + out.stderr[-2000:]
    r = json.loads(lines[-1])
    single = [ln for ln in out.stdout.splitlines() if ln.startswith("single-GPU:")][-1].split()
    assert r["status"] == "Solved" and r["repeat_identical"]
    assert r["pres_cpu"] < 1.5e-4 and r["dres_cpu"] < 1.5e-4
    assert int(single[3]) == r["admm"] and float(single[-1]) < 1e-9    # same iterations, x agrees
