"""The real-data branch of the loaders (reference utils.py:141-184: ImageFolder + RandomResizedCrop / Resize+CenterCrop +
ToTensor + Normalize) on a generated 2-class, 16-image JPEG folder -- there is no ImageNet offline, but the code path
(torchvision.datasets.ImageFolder, the transforms, the DistributedSampler, B200.INPUT_UINT8's PILToTensor tail and
the CLI with MODEL.DUMMY_INPUT False) must still have run somewhere."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture
def image_root(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(0)
    for split in ("train", "val"):
        for cls in ("n01", "n02"):
            d = tmp_path / "data" / split / cls
            d.mkdir(parents=True)
            for i in range(4):
                h, w = int(rng.integers(80, 140)), int(rng.integers(80, 140))
                Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(d / f"img_{i}.jpg", quality=90)
    return str(tmp_path / "data")


def test_imagefolder_loaders_fp32_and_uint8(fresh_cfg, image_root):
    from distribuuuu_b200 import utils
    cfg = fresh_cfg
    cfg.MODEL.DUMMY_INPUT = False
    cfg.TRAIN.DATASET, cfg.TRAIN.BATCH_SIZE, cfg.TEST.BATCH_SIZE, cfg.TRAIN.WORKERS = image_root, 4, 4, 0
    cfg.TRAIN.IM_SIZE, cfg.TEST.IM_SIZE = 64, 72
    train, val = utils.construct_train_loader(), utils.construct_val_loader()
    assert len(train.dataset) == 8 and len(val.dataset) == 8 and train.dataset.classes == ["n01", "n02"]
    train.sampler.set_epoch(0)
    x, y = next(iter(train))
    assert x.shape == (4, 3, 64, 64) and x.dtype == torch.float32 and y.dtype == torch.int64
    assert abs(float(x.mean())) < 1.5 and float(x.std()) > 0.5            # normalised, not raw [0,1]
    xv, yv = next(iter(val))
    assert xv.shape == (4, 3, 224, 224)                                  # reference utils.py:160 crops 224 regardless
    assert sorted(torch.cat([b[1] for b in val]).tolist()) == [0] * 4 + [1] * 4
    # B200.INPUT_UINT8: raw pixels leave the loader, the same images normalise to the same tensors on the device side
    cfg.B200.INPUT_UINT8 = True
    val8 = utils.construct_val_loader()
    x8, y8 = next(iter(val8))
    assert x8.dtype == torch.uint8 and x8.shape == (4, 3, 224, 224) and torch.equal(y8, yv)
    assert torch.allclose(utils.normalize_uint8(x8), xv, atol=1e-5)


def test_cli_trains_on_an_image_folder(tmp_path, image_root, free_port):
    out = str(tmp_path / "exp")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port), os.path.join(ROOT, "train_net.py"), "--cfg", os.path.join(ROOT, "config", "resnet18.yaml"),
           "MODEL.DUMMY_INPUT", "False", "TRAIN.DATASET", image_root, "MODEL.NUM_CLASSES", "2", "TRAIN.BATCH_SIZE", "2",
           "TEST.BATCH_SIZE", "2", "TRAIN.IM_SIZE", "64", "TEST.IM_SIZE", "72", "TRAIN.WORKERS", "0", "OPTIM.MAX_EPOCH", "1",
           "TRAIN.TOPK", "1", "OUT_DIR", out, "B200.DEVICE", "cpu", "TRAIN.PRINT_FREQ", "1"]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "TRAIN:  [1]" in r.stderr and "ACCURACY: TOP1" in r.stderr
    assert os.path.exists(os.path.join(out, "checkpoints", "ckpt_ep_001.pth.tar"))
