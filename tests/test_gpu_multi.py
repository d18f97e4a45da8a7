"""Peer-memory collectives on >= 2 GPUs: fused all-reduce+SGD, SyncBN exchange, engine parity (tools/multigpu_check.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def test_peer_memory_paths_two_ranks(free_port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port), os.path.join(ROOT, "tools", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    assert "PASS allreduce_sgd" in r.stdout and "PASS syncbn" in r.stdout and "PASS engine" in r.stdout
