"""Peer-memory collectives on >= 2 GPUs: fused all-reduce+SGD, SyncBN exchange, fp32-master consistency, engine parity
(tools/multigpu_check.py) and the train -> resume -> test CLI under torchrun with one rank per GPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(nproc, port, script, *args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_peer_memory_paths(free_port, nproc):
    if _n_gpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    r = _torchrun(nproc, free_port, "tools/multigpu_check.py")
    assert r.returncode == 0, (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    for name in ("allreduce_sgd", "syncbn", "fp32_masters", "engine", "graph_replay"):
        assert f"PASS {name}" in r.stdout, r.stdout[-3000:]


def test_cli_train_resume_test_all_gpus(tmp_path, free_port):
    """torchrun --nproc-per-node=<all GPUs> train_net.py (native engine, SyncBN, peer-memory all-reduce) -> auto-resume
    -> test_net.py on the best checkpoint."""
    nproc = min(_n_gpus(), 8)
    out = str(tmp_path / "exp")
    common = ["--cfg", os.path.join(ROOT, "config", "resnet50.yaml"), "MODEL.DUMMY_INPUT", "True", "MODEL.SYNCBN", "True",
              "B200.DUMMY_ON_DEVICE", "True", "TRAIN.BATCH_SIZE", "16", "TEST.BATCH_SIZE", "16", "B200.DUMMY_LEN",
              str(64 * nproc), "TRAIN.WORKERS", "0", "OUT_DIR", out, "B200.MAX_ITERS", "3", "TRAIN.PRINT_FREQ", "1"]
    r = _torchrun(nproc, free_port, "train_net.py", *common, "OPTIM.MAX_EPOCH", "1")
    assert r.returncode == 0, r.stderr[-3000:]
    assert os.path.exists(os.path.join(out, "checkpoints", "ckpt_ep_001.pth.tar")), r.stderr[-2000:]
    assert "TRAIN:  [1]" in r.stderr and "ACCURACY: TOP1" in r.stderr
    r2 = _torchrun(nproc, free_port + 1, "train_net.py", *common, "OPTIM.MAX_EPOCH", "2")
    assert r2.returncode == 0, r2.stderr[-3000:]
    assert "LOADED" in r2.stderr and "TRAIN:  [2]" in r2.stderr and "TRAIN:  [1]" not in r2.stderr
    r3 = _torchrun(nproc, free_port + 2, "test_net.py", *common, "MODEL.WEIGHTS", os.path.join(out, "best.pth.tar"))
    assert r3.returncode == 0, r3.stderr[-3000:]
    assert "ACCURACY: TOP1" in r3.stderr
