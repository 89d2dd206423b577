"""GPU, >= 2 devices, NCCL: the N>1 path on hardware (skipped on a 1-GPU box; `gpurun --gpus 2` runs it).
The checks live in tests/mgpu_worker.py (one process per GPU, launched with torch.distributed.run)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_solve_global_stop_and_gradient_allreduce_nccl():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("MGPU_")]
    assert out.returncode == 0 and lines and lines[-1].startswith("MGPU_OK"), out.stdout[-3000:] + out.stderr[-3000:]
    rep = json.loads(lines[-1].split(" ", 1)[1])
    assert rep["per_shard_equal"] and rep["global_stop_equal"] and rep["grad_allreduce"]
    # the scenario must actually separate the two stop rules, otherwise the global-stop check proves nothing
    assert rep["global_differs_from_some_shard"], rep
