"""train_net.py / test_net.py end to end on the native engine (synthetic data, a few iterations)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(script, port, out_dir, extra, nproc=1):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, script), "--cfg",
           os.path.join(ROOT, "config", "resnet50.yaml"), "MODEL.DUMMY_INPUT", "True", "B200.DUMMY_ON_DEVICE", "True",
           "TRAIN.BATCH_SIZE", "16", "TEST.BATCH_SIZE", "16", "B200.DUMMY_LEN", "64", "TRAIN.WORKERS", "0", "OUT_DIR", out_dir,
           "B200.MAX_ITERS", "3", "TRAIN.PRINT_FREQ", "1"] + extra
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)


def test_train_resume_eval_native_engine(tmp_path, free_port):
    out = str(tmp_path / "exp")
    r = _run("train_net.py", free_port, out, ["OPTIM.MAX_EPOCH", "1"])
    assert r.returncode == 0, r.stderr[-3000:]
    assert os.path.exists(os.path.join(out, "checkpoints", "ckpt_ep_001.pth.tar")), r.stderr[-2000:]
    assert "TRAIN:  [1]" in r.stderr and "ACCURACY: TOP1" in r.stderr
    r2 = _run("train_net.py", free_port + 1, out, ["OPTIM.MAX_EPOCH", "2"])
    assert r2.returncode == 0, r2.stderr[-3000:]
    assert "LOADED" in r2.stderr and "TRAIN:  [2]" in r2.stderr and "TRAIN:  [1]" not in r2.stderr
    r3 = _run("test_net.py", free_port + 2, out, ["MODEL.WEIGHTS", os.path.join(out, "best.pth.tar")])
    assert r3.returncode == 0, r3.stderr[-3000:]
    assert "ACCURACY: TOP1" in r3.stderr


def test_train_with_captured_step(tmp_path, free_port):
    """B200.CUDA_GRAPH: three eager steps, then the step is captured and replayed (two statistics parities -> two graphs);
    the run trains, checkpoints and evaluates exactly as without it."""
    out = str(tmp_path / "exp_graph")
    r = _run("train_net.py", free_port, out, ["OPTIM.MAX_EPOCH", "1", "B200.CUDA_GRAPH", "True", "B200.MAX_ITERS", "10",
                                              "B200.DUMMY_LEN", "256"])
    assert r.returncode == 0, r.stderr[-3000:]
    assert "captured the training step in a CUDA graph" in r.stderr, r.stderr[-3000:]
    assert os.path.exists(os.path.join(out, "checkpoints", "ckpt_ep_001.pth.tar")) and "ACCURACY: TOP1" in r.stderr
