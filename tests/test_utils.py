"""LR policy, meters, accuracy, checkpoint layout (reference utils.py:199-410)."""
import math
import os

import pytest
import torch

from distribuuuu_b200 import utils
from distribuuuu_b200.models import build_model


def test_cosine_lr_with_warmup(fresh_cfg):
    c = fresh_cfg
    c.OPTIM.MAX_EPOCH, c.OPTIM.BASE_LR, c.OPTIM.WARMUP_EPOCHS, c.OPTIM.WARMUP_FACTOR = 100, 0.2, 5, 0.1
    for e in range(0, 100, 7):
        want = 0.5 * (1 + math.cos(math.pi * e / 100)) * 0.2
        if e < 5:
            a = e / 5
            want *= 0.1 * (1 - a) + a
        assert utils.get_epoch_lr(e) == pytest.approx(want)
    assert utils.get_epoch_lr(0) == pytest.approx(0.02)
    c.OPTIM.MIN_LR = 0.25  # a fraction of BASE_LR, not an absolute value
    assert utils.get_epoch_lr(100) == pytest.approx(0.25 * 0.2)


def test_steps_lr(fresh_cfg):
    c = fresh_cfg
    c.OPTIM.LR_POLICY, c.OPTIM.STEPS, c.OPTIM.LR_MULT, c.OPTIM.WARMUP_EPOCHS = "steps", [0, 30, 60, 90], 0.1, 0
    assert [round(utils.get_epoch_lr(e), 6) for e in (0, 29, 30, 61, 95)] == [0.2, 0.2, 0.02, 0.002, 0.0002]
    c.OPTIM.LR_POLICY = "exp"
    with pytest.raises(AssertionError):
        utils.get_epoch_lr(1)


def test_set_lr():
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    utils.set_lr(opt, 0.5)
    assert opt.param_groups[0]["lr"] == 0.5


def test_meters_format(fresh_cfg):
    m = utils.AverageMeter("Loss", ":6.4f")
    m.update(2.0, 2)
    m.update(1.0, 2)
    assert m.avg == 1.5 and str(m) == "Loss 1.0000 (1.5000)"
    meters = utils.construct_meters()
    assert [x.name for x in meters] == ["Time", "Data", "Loss", "Acc@1", "Acc@5"]
    p = utils.ProgressMeter(626, meters, prefix="TRAIN:  [1]")
    assert p._batch_fmt.format(7) == "[  7/626]"


def test_accuracy_matches_definition():
    torch.manual_seed(0)
    out = torch.randn(64, 10)
    tgt = torch.randint(0, 10, (64,))
    a1, a5 = utils.accuracy(out, tgt, topk=(1, 5))
    top5 = out.topk(5, 1).indices
    assert a1.shape == (1,)
    assert a1.item() == pytest.approx(100.0 * (out.argmax(1) == tgt).float().mean().item())
    assert a5.item() == pytest.approx(100.0 * (top5 == tgt[:, None]).any(1).float().mean().item())


def test_device_metrics_single_process():
    dm = utils.DeviceMetrics(torch.device("cpu"))
    dm.update(torch.tensor(2.0), torch.tensor(3), torch.tensor(4), 4)
    dm.update(torch.tensor(1.0), torch.tensor(1), torch.tensor(4), 4)
    loss, t1, tk, n = dm.flush()
    assert (loss, t1, tk, n) == (1.5, 50.0, 100.0, 8)


def test_checkpoint_paths_and_roundtrip(fresh_cfg, tmp_path):
    fresh_cfg.OUT_DIR = str(tmp_path)
    assert not utils.has_checkpoint()
    assert utils.get_checkpoint(7).endswith(os.path.join("checkpoints", "ckpt_ep_007.pth.tar"))
    net = build_model("resnet18", num_classes=10)
    opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9, nesterov=True, weight_decay=5e-5)
    net(torch.randn(2, 3, 32, 32)).sum().backward()
    opt.step()
    path = utils.save_checkpoint(net, opt, epoch=4, best_acc1=12.5, best=True)
    assert path == utils.get_checkpoint(5) and os.path.exists(tmp_path / "best.pth.tar")
    utils.save_checkpoint(net, opt, epoch=10, best_acc1=12.5, best=False)
    assert utils.get_last_checkpoint() == utils.get_checkpoint(11)
    ckpt = torch.load(path, weights_only=False)
    assert set(ckpt) == {"epoch", "state_dict", "optimizer", "best_acc1"}
    assert not any(k.startswith("module.") for k in ckpt["state_dict"])
    assert "momentum_buffer" in ckpt["optimizer"]["state"][0]

    net2 = build_model("resnet18", num_classes=10)
    opt2 = torch.optim.SGD(net2.parameters(), lr=0.1, momentum=0.9, nesterov=True)
    start, best = utils.load_checkpoint(path, net2, opt2)
    assert (start, best) == (5, 12.5)
    for a, b in zip(net.state_dict().values(), net2.state_dict().values()):
        assert torch.equal(a, b)
    # bare state_dict (best.pth.tar) loads too, epoch/best stay 0
    assert utils.load_checkpoint(str(tmp_path / "best.pth.tar"), net2, opt2) == (0, 0)
    # weights only when no optimizer is passed
    assert utils.load_checkpoint(path, net2) == (0, 0)


def test_torchvision_checkpoint_interop():
    """Keys are torchvision's: a torchvision resnet50 state_dict loads strictly."""
    tv = pytest.importorskip("torchvision")
    ours = build_model("resnet50")
    theirs = tv.models.resnet50()
    ours.load_state_dict(theirs.state_dict(), strict=True)
    x = torch.randn(1, 3, 64, 64)
    ours.eval(), theirs.eval()
    with torch.no_grad():
        assert torch.allclose(ours(x), theirs(x), atol=1e-5)
    d_ours, d_theirs = build_model("densenet121"), tv.models.densenet121()
    d_ours.load_state_dict(d_theirs.state_dict(), strict=True)
    d_ours.eval(), d_theirs.eval()
    with torch.no_grad():
        assert torch.allclose(d_ours(x), d_theirs(x), atol=1e-5)


def test_dummy_dataset_and_loaders(fresh_cfg):
    fresh_cfg.MODEL.DUMMY_INPUT = True
    fresh_cfg.B200.DUMMY_LEN = 40
    fresh_cfg.TRAIN.BATCH_SIZE, fresh_cfg.TEST.BATCH_SIZE, fresh_cfg.TRAIN.WORKERS = 16, 16, 0
    fresh_cfg.TRAIN.IM_SIZE = 32
    tl, vl = utils.construct_train_loader(), utils.construct_val_loader()
    assert len(tl) == 2 and len(vl) == 3  # drop_last vs keep
    x, y = next(iter(tl))
    assert x.shape == (16, 3, 32, 32) and int(y.sum()) == 0
    pf = utils.PinnedPrefetcher(tl, torch.device("cpu"))
    assert sum(1 for _ in pf) == 2


def test_uint8_input_pipeline(fresh_cfg):
    """B200.INPUT_UINT8: loaders hand out raw uint8 pixels; the torch engine normalises them like ToTensor+Normalize."""
    from distribuuuu_b200 import models
    from distribuuuu_b200.trainer import TorchEngine
    fresh_cfg.MODEL.DUMMY_INPUT = True
    fresh_cfg.B200.INPUT_UINT8 = True
    fresh_cfg.B200.DUMMY_LEN = 16
    fresh_cfg.TRAIN.BATCH_SIZE, fresh_cfg.TEST.BATCH_SIZE, fresh_cfg.TRAIN.WORKERS = 8, 8, 0
    fresh_cfg.TRAIN.IM_SIZE = 32
    x, y = next(iter(utils.construct_train_loader()))
    assert x.dtype == torch.uint8 and x.shape == (8, 3, 32, 32)
    want = (x.float() / 255.0 - torch.tensor(utils.IMAGENET_MEAN).view(1, 3, 1, 1)) / torch.tensor(utils.IMAGENET_STD).view(1, 3, 1, 1)
    assert torch.allclose(utils.normalize_uint8(x), want, atol=1e-6)
    torch.manual_seed(0)
    eng = TorchEngine(models.build_model("resnet18", num_classes=10))
    eng.eval()
    t = torch.zeros(8, dtype=torch.long)
    la, _, _ = eng.eval_step(x, t, 5)
    lb, _, _ = eng.eval_step(want, t, 5)
    assert abs(float(la) - float(lb)) < 1e-5
    # the image-folder tail keeps PIL -> uint8 CHW without Normalize
    import torchvision.transforms as T
    from distribuuuu_b200.utils import data as D
    tail = D._tail_transforms(T)
    assert len(tail) == 1 and isinstance(tail[0], T.PILToTensor)


def test_step_watchdog_fires_and_recovers():
    import time
    fired = []
    wd = utils.StepWatchdog(0.15, abort=False, on_timeout=fired.append, name="t").start()
    for _ in range(5):          # regular ticks: silent
        time.sleep(0.03)
        wd.tick()
    assert wd.fired == 0
    time.sleep(0.5)             # stall: fires (and would have dumped the stacks / aborted in a real job)
    wd.stop()
    assert wd.fired >= 1 and fired and fired[0] >= 0.15
    off = utils.StepWatchdog(0, name="off").start()   # disabled: no thread at all
    assert off._thread is None
    off.stop()


def test_heartbeat_files_and_stale_detection(tmp_path):
    import time
    hb0, hb1 = utils.Heartbeat(str(tmp_path), 0, 2), utils.Heartbeat(str(tmp_path), 1, 2)
    hb0.beat(0, 1)               # not a multiple of freq: skipped
    assert utils.read_heartbeats(str(tmp_path)) == []
    hb0.beat(0, 2)
    hb1.beat(0, 2)
    beats = utils.read_heartbeats(str(tmp_path))
    assert [b["rank"] for b in beats] == [0, 1] and beats[0]["iter"] == 2
    assert utils.stale_ranks(str(tmp_path), 60.0) == []
    assert utils.stale_ranks(str(tmp_path), 60.0, now=time.time() + 120) == [0, 1]
    utils.Heartbeat(str(tmp_path), 2, 0).beat(0, 2)   # disabled
    assert len(utils.read_heartbeats(str(tmp_path))) == 2
