"""Multi-GPU parity (``-m gpu``; skipped on boxes with a single GPU): launches tools/dist_check.py under
torchrun with one rank per GPU over NCCL and requires every check to pass on every rank."""

import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_dist_check_under_torchrun():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "dist_check.py")]  # fmt: skip
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    tail = (res.stdout + res.stderr)[-4000:]
    assert res.returncode == 0, tail
    assert f"dist_check: OK on {nproc} ranks" in res.stdout, tail
