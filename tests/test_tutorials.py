"""The tutorial ladder runs end to end on CPU with synthetic data (reference tutorial/*.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ["--synthetic", "--device", "cpu", "--epochs", "1", "--max-iters", "2", "--batch-size", "8", "--print-freq", "1"]


def _run(cmd, cwd, timeout=600):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=cwd)


def test_single_process_script(tmp_path):
    r = _run([sys.executable, os.path.join(ROOT, "tutorial", "snsc.py")] + COMMON, str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Training Finished" in r.stdout and "loss:" in r.stdout


@pytest.mark.parametrize("script", ["mnmc_ddp_launch.py", "imagenet.py"])
def test_torchrun_scripts(script, tmp_path, free_port):
    extra = ["--ckpt", str(tmp_path / "t.pth.tar")] if script == "imagenet.py" else []
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port), os.path.join(ROOT, "tutorial", script)] + COMMON + extra
    r = _run(cmd, str(tmp_path))
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert "loss:" in r.stdout
    if extra:
        assert os.path.exists(extra[1])


def test_spawn_script(tmp_path, free_port):
    cmd = [sys.executable, os.path.join(ROOT, "tutorial", "mnmc_ddp_mp.py"), "--nproc-per-node", "2", "--port", str(free_port)] + COMMON
    r = _run(cmd, str(tmp_path))
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert "loss:" in r.stdout
