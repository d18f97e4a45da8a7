"""2-process gloo tests of the plumbing (BASELINE.json config[0]: resnet18, world 2, CPU, synthetic)."""
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _spawn(fn, world, port, *args):
    mp.spawn(_entry, args=(fn, world, port, args), nprocs=world, join=True)


def _entry(rank, fn, world, port, args):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from distribuuuu_b200 import config, utils
    config.reset_cfg()
    config.cfg.B200.DEVICE = "cpu"
    utils.setup_distributed()
    try:
        fn(rank, world, *args)
    finally:
        utils.shutdown()


def _check_scaled_all_reduce(rank, world):
    from distribuuuu_b200 import utils
    a, b = torch.tensor(float(rank + 1)), torch.tensor([10.0 * (rank + 1)])
    utils.scaled_all_reduce([a, b])
    assert a.item() == pytest.approx(1.5) and b.item() == pytest.approx(15.0)
    dm = utils.DeviceMetrics(torch.device("cpu"))
    dm.update(torch.tensor(float(rank)), torch.tensor(rank), torch.tensor(2), 2)
    loss, t1, tk, n = dm.flush()
    assert (loss, t1, tk, n) == (0.5, 25.0, 100.0, 4)


def _check_ddp_matches_large_batch(rank, world):
    """Averaged-gradient data parallel == single process on the concatenated batch."""
    from distribuuuu_b200.models import build_model
    from distribuuuu_b200.trainer import TorchEngine
    torch.manual_seed(100)  # == rank 0's init, which the engine broadcasts
    ref = build_model("resnet18", num_classes=10)
    for m in ref.modules():  # BN batch statistics differ per shard; use eval-mode BN for exact parity
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    torch.manual_seed(100 + rank)  # different init per rank: the engine must broadcast rank 0's
    net = build_model("resnet18", num_classes=10)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    eng = TorchEngine(net, bucket_cap_mb=4)
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(8, 3, 32, 32, generator=g), torch.randint(0, 10, (8,), generator=g)
    opt = torch.optim.SGD(eng.parameters(), lr=0.1, momentum=0.9, nesterov=True, weight_decay=5e-5)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, nesterov=True, weight_decay=5e-5)
    for _ in range(2):
        eng.train_step(x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4], opt, 5)
        ropt.zero_grad()
        torch.nn.functional.cross_entropy(ref(x), y).backward()
        ropt.step()
    for (n, a), b in zip(net.named_parameters(), ref.parameters()):
        assert torch.allclose(a, b, atol=2e-5), n


def _check_syncbn_matches_global_bn(rank, world):
    from distribuuuu_b200.parallel import SyncBatchNorm
    torch.manual_seed(0)
    bn = torch.nn.BatchNorm2d(6)
    sbn = SyncBatchNorm.convert_sync_batchnorm(torch.nn.Sequential(torch.nn.BatchNorm2d(6)))[0]
    assert isinstance(sbn, SyncBatchNorm)
    with torch.no_grad():
        for m in (bn, sbn):
            m.weight.copy_(torch.linspace(0.5, 1.5, 6))
            m.bias.copy_(torch.linspace(-1, 1, 6))
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 6, 5, 5, generator=g)
    w = torch.randn(8, 6, 5, 5, generator=g)
    xf = x.clone().requires_grad_(True)
    ref_out = bn(xf)
    (ref_out * w).sum().backward()
    xs = x[rank * 4:(rank + 1) * 4].clone().requires_grad_(True)
    out = sbn(xs)
    (out * w[rank * 4:(rank + 1) * 4]).sum().backward()
    assert torch.allclose(out, ref_out[rank * 4:(rank + 1) * 4].detach(), atol=1e-5)
    assert torch.allclose(xs.grad, xf.grad[rank * 4:(rank + 1) * 4], atol=1e-5)
    assert torch.allclose(sbn.running_mean, bn.running_mean, atol=1e-6)
    assert torch.allclose(sbn.running_var, bn.running_var, atol=1e-5)
    gw = sbn.weight.grad.clone()
    dist.all_reduce(gw)
    assert torch.allclose(gw, bn.weight.grad, atol=1e-4)


def _check_native_engine_bucket_schedule(rank, world):
    """The native engine's multi-rank host logic on gloo + the emulated kernel module (tests/fake_kernels.py):
    rank-0 weights are broadcast, every bucket's fused update is launched from inside backward as soon as its last
    gradient exists (roughly reverse registration order, identical on all ranks), left-over buckets are flushed,
    and the result equals one process
    training on the concatenated batch (B200.COMM=nccl code path: library all-reduce + local fused update)."""
    import contextlib
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_kernels import FakeKernels
    from distribuuuu_b200.models import build_model
    from distribuuuu_b200.ops import build
    from distribuuuu_b200.parallel import native_engine

    class _Dummy:                       # stands in for CUDA streams / events
        def __init__(self, *a, **k): pass
        def wait_event(self, *a): pass
        def wait_stream(self, *a): pass
        def record(self, *a): pass

    fake = FakeKernels()
    build.load = lambda *a, **k: fake
    torch.cuda.Stream = torch.cuda.Event = _Dummy
    torch.cuda.current_stream = lambda *a, **k: _Dummy()
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.synchronize = lambda *a, **k: None

    def make(seed):
        torch.manual_seed(seed)
        net = build_model("resnet18", num_classes=10)
        for m in net.modules():         # per-shard batch statistics differ from the global ones: eval-mode BN
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
        return net

    ref = make(100)
    eng = native_engine.NativeEngine(make(100 + rank), torch.device("cpu"), comm="nccl", bucket_cap_mb=4)
    assert eng.world == 2 and eng.comm_mode == "nccl" and len(eng.buckets) >= 3
    launched = []
    real_launch = eng._launch_bucket
    eng._launch_bucket = lambda b: (launched.append((eng.buckets.index(b), eng._in_train_step)), real_launch(b))
    opt = eng.make_optimizer(lr=0.05, momentum=0.9, dampening=0.0, weight_decay=5e-5, nesterov=True)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, nesterov=True, weight_decay=5e-5)
    g = torch.Generator().manual_seed(5)
    x, y = torch.randn(8, 3, 32, 32, generator=g), torch.randint(0, 10, (8,), generator=g)
    for _ in range(2):
        launched.clear()
        eng.train_step(x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4], opt, 5)
        ropt.zero_grad()
        torch.nn.functional.cross_entropy(ref(x), y).backward()
        ropt.step()
        order = [i for i, _ in launched]
        assert sorted(order) == list(range(len(eng.buckets))), order                    # every bucket exactly once
        assert order[0] == 0 and order != sorted(order)   # classifier first; projection shortcuts finish out of order
        orders = [None] * world
        dist.all_gather_object(orders, order)
        assert orders[0] == orders[1]                      # collectives are issued in the same order on every rank
        assert sum(1 for _, inside in launched if inside) >= len(eng.buckets) - 1       # overlapped with backward
    mine = eng.flat_master.clone()
    other = mine.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(mine, other), float((mine - other).abs().max())                  # ranks stay bit-identical
    dot = na = nb = 0.0
    init = make(100)
    for (n, a), b, i in zip(eng.module.named_parameters(), ref.parameters(), init.parameters()):
        da, db = (a.detach() - i.detach()).flatten(), (b.detach() - i.detach()).flatten()
        dot, na, nb = dot + float(da @ db), na + float(da @ da), nb + float(db @ db)
    assert dot / (na ** 0.5 * nb ** 0.5) > 0.95 and 0.9 < (na / nb) ** 0.5 < 1.1      # bf16 compute vs fp32 reference


@pytest.mark.parametrize("fn", [_check_scaled_all_reduce, _check_ddp_matches_large_batch,
                                _check_syncbn_matches_global_bn, _check_native_engine_bucket_schedule])
def test_two_rank_gloo(fn, free_port):
    _spawn(fn, 2, free_port)


def _run_cli(script, port, out_dir, extra, nproc=2):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, script),
           "--cfg", os.path.join(ROOT, "config", "resnet18.yaml"), "MODEL.DUMMY_INPUT", "True",
           "TRAIN.BATCH_SIZE", "4", "TEST.BATCH_SIZE", "4", "TRAIN.IM_SIZE", "32", "B200.DUMMY_LEN", "16",
           "TRAIN.WORKERS", "0", "OUT_DIR", out_dir, "B200.DEVICE", "cpu", "B200.MAX_ITERS", "2",
           "MODEL.NUM_CLASSES", "10"] + extra
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_train_resume_test_cli_end_to_end(tmp_path, free_port):
    """train_net.py (2 ranks, gloo) -> checkpoints -> auto-resume -> test_net.py."""
    out = str(tmp_path / "exp")
    r = _run_cli("train_net.py", free_port, out, ["OPTIM.MAX_EPOCH", "1", "MODEL.SYNCBN", "True"])
    assert r.returncode == 0, r.stderr[-2000:]
    assert os.path.exists(os.path.join(out, "checkpoints", "ckpt_ep_001.pth.tar"))
    assert os.path.exists(os.path.join(out, "config.yaml")) and os.path.exists(os.path.join(out, "best.pth.tar"))
    assert "ACCURACY: TOP1" in r.stderr and "TRAIN:  [1]" in r.stderr

    # second epoch: auto-resume; also exercises the uint8 input path, the heartbeat files and the (silent) watchdog
    r2 = _run_cli("train_net.py", free_port + 1, out, ["OPTIM.MAX_EPOCH", "2", "B200.INPUT_UINT8", "True",
                                                       "B200.HEARTBEAT_FREQ", "1", "B200.WATCHDOG_S", "300"])
    assert r2.returncode == 0, r2.stderr[-2000:]
    assert "LOADED" in r2.stderr and "TRAIN:  [2]" in r2.stderr and "TRAIN:  [1]" not in r2.stderr
    assert os.path.exists(os.path.join(out, "checkpoints", "ckpt_ep_002.pth.tar"))
    from distribuuuu_b200 import utils
    beats = utils.read_heartbeats(out)
    assert [b["rank"] for b in beats] == [0, 1] and all(b["epoch"] == 1 and b["iter"] == 2 for b in beats)

    r3 = _run_cli("test_net.py", free_port + 2, out, ["MODEL.WEIGHTS", os.path.join(out, "best.pth.tar")])
    assert r3.returncode == 0, r3.stderr[-2000:]
    assert "ACCURACY: TOP1" in r3.stderr


def test_slurm_env_bootstrap(monkeypatch, free_port):
    """Slurm branch derives the torchrun contract (reference utils.py:26-40)."""
    from distribuuuu_b200 import config, utils
    config.reset_cfg()
    config.cfg.B200.DEVICE = "cpu"
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("SLURM_JOB_ID", "1")
    monkeypatch.setenv("SLURM_PROCID", "0")
    monkeypatch.setenv("SLURM_NTASKS", "1")
    monkeypatch.setenv("SLURM_NODELIST", "localhost")
    monkeypatch.setenv("MASTER_ADDR", "127.0.0.1")
    utils.setup_distributed(port=free_port)
    try:
        assert os.environ["RANK"] == "0" and os.environ["WORLD_SIZE"] == "1"
        assert os.environ["MASTER_PORT"] == str(free_port) and os.environ["LOCAL_RANK"] == "0"
        assert dist.get_backend() == "gloo"
    finally:
        utils.shutdown()
        config.reset_cfg()
