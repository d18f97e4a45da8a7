"""The native engine's host logic, end to end on CPU: the real ``NativeEngine`` / ``ops.native`` code drives an
emulated kernel module (tests/fake_kernels.py) and must reproduce the fp32 torch path step for step -- forward, the
gradient routing between block branches (mailboxes instead of add kernels), strided-dgrad decompositions, BN
statistic slots, flat-buffer views and the fused optimizer step."""
import copy
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fake_kernels import FakeKernels  # noqa: E402

from distribuuuu_b200.models import build_model  # noqa: E402
from distribuuuu_b200.ops import build, functional as Fn  # noqa: E402
from distribuuuu_b200.parallel import native_engine  # noqa: E402


@pytest.fixture
def cpu_engine(monkeypatch):
    fake = FakeKernels()
    monkeypatch.setattr(build, "load", lambda *a, **k: fake)
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)

    def make(arch, **kw):
        torch.manual_seed(0)
        if arch == "bottleneck_tiny":   # every Bottleneck code path (identity + both projection kinds) at depth 5,
            from distribuuuu_b200.models.resnet import Bottleneck, ResNet   # shallow enough for a per-tensor check
            net = ResNet(Bottleneck, [2, 1, 1, 1], **kw)
        else:
            net = build_model(arch, **kw)
        ref = copy.deepcopy(net)
        eng = native_engine.NativeEngine(net, torch.device("cpu"))
        return eng, ref, fake
    return make


def _train(eng, ref, steps, batch, size, classes, lr=0.02):
    opt = eng.make_optimizer(lr=lr, momentum=0.9, dampening=0.0, weight_decay=5e-5, nesterov=True)
    ropt = torch.optim.SGD(ref.parameters(), lr=lr, momentum=0.9, weight_decay=5e-5, nesterov=True)
    eng.train(), ref.train()
    g = torch.Generator().manual_seed(1)
    losses, first_update = [], None
    init = {k: v.clone() for k, v in ref.state_dict().items()}
    for step in range(steps):
        x = torch.randn(batch, 3, size, size, generator=g)
        y = torch.randint(0, classes, (batch,), generator=g)
        la, h1, hk = eng.train_step(x, y, opt, 5)
        lb, _, _ = Fn.cross_entropy_topk(ref(x), y, 5)
        ropt.zero_grad()
        lb.backward()
        ropt.step()
        losses.append((float(la.detach()), float(lb.detach())))
        if step == 0:   # per-tensor agreement of the very first update (later steps compound bf16 noise)
            sa, sb = eng.module.state_dict(), ref.state_dict()
            first_update = {}
            for k, v in sb.items():
                if v.dtype.is_floating_point and "running" not in k:
                    da, db = (sa[k].float() - init[k]).flatten(), (v - init[k]).flatten()
                    first_update[k] = (da, db)
                elif k.endswith("running_mean") or k.endswith("running_var"):
                    first_update[k] = (sa[k].clone(), v.clone())
    return losses, first_update


@pytest.mark.parametrize("arch,size", [("resnet18", 64), ("bottleneck_tiny", 64)])
def test_training_steps_match_the_torch_path(cpu_engine, arch, size):
    eng, ref, fake = cpu_engine(arch, num_classes=16)
    losses, upd = _train(eng, ref, steps=3, batch=8, size=size, classes=16)
    for la, lb in losses:
        assert abs(la - lb) / max(abs(lb), 1e-3) < 0.05, losses          # bf16 activations vs fp32 reference
    # The first fused update points where torch.optim.SGD's does, tensor by tensor.  bf16 activations make the
    # early layers noisier (cos ~0.9) than the classifier (cos ~1.0); a mis-routed gradient gives cos ~0 or a norm
    # that is off by the fan-in, which is what this guards against.
    dot = na = nb = 0.0
    for k, (da, db) in upd.items():
        if "running" in k:              # running statistics after the first step (same inputs, same weights)
            assert torch.allclose(da, db, atol=2e-2, rtol=2e-2), k
            continue
        cos = float(torch.dot(da, db) / (da.norm() * db.norm() + 1e-20))
        ratio = float(da.norm() / (db.norm() + 1e-20))
        assert cos > 0.7 and 0.7 < ratio < 1.4, (k, cos, ratio)
        dot, na, nb = dot + float(torch.dot(da, db)), na + float(da.norm() ** 2), nb + float(db.norm() ** 2)
    assert dot / (na ** 0.5 * nb ** 0.5) > 0.88 and 0.93 < (na / nb) ** 0.5 < 1.07
    sd_a, sd_b = eng.module.state_dict(), ref.state_dict()
    for k, v in sd_b.items():           # every BN layer advanced its counter once per step
        if k.endswith("num_batches_tracked"):
            assert int(sd_a[k]) == int(v) == 3
    # the gradient buffers were consumed (zeroed) by the fused update, bf16 compute weights follow the masters
    assert float(eng.flat_grad.abs().max()) == 0.0
    assert float((eng.flat_w16.float() - eng.flat_master).abs().max()) < 2e-2
    # gradient fan-in went through the mailboxes: residual-branch gradients rode along in a dgrad epilogue
    assert fake.calls.get("conv_dgrad+addend", 0) > 0 and fake.calls.get("bn_backward+mask", 0) > 0
    # the stem's BN + ReLU + max-pool ran as the fused tail (one forward, one backward call per step), not as three ops
    assert fake.calls.get("bn_relu_pool_fwd", 0) == 3 == fake.calls.get("bn_relu_pool_bwd", 0)
    assert fake.calls.get("maxpool_fwd", 0) == 0


@pytest.mark.parametrize("arch,size,batch", [("efficientnet_b0", 128, 16), ("densenet121", 64, 8), ("regnety_160", 64, 4),
                                             ("resnext50_32x4d", 64, 8), ("botnet50", 224, 2)])
def test_other_model_families_step_like_the_torch_path(cpu_engine, arch, size, batch):
    """Depthwise + fused SE + SiLU (EfficientNet), pre-activation BN + concat + avg-pool (DenseNet), wide grouped
    convs + fused SE (RegNetY), thin groups as 64-channel block-diagonal groups (ResNeXt), relative-position
    attention between native projections (BoTNet, fixed 224x224 input).  No call may leave the kernel module for an
    ATen / library op (``NativeOps.fallbacks`` stays empty)."""
    eng, ref, fake = cpu_engine(arch, num_classes=16)
    # same batch / resolution / tolerance as the GPU parity checks (tools/gpu_selftest.py): small BN sample counts
    # make these nets sensitive to bf16 rounding, so the bound is on the loss trajectory, not per tensor
    losses, _ = _train(eng, ref, steps=2, batch=batch, size=size, classes=16, lr=0.005)
    for la, lb in losses:
        assert abs(la - lb) / max(abs(lb), 1e-3) < 0.15, losses
    assert float(eng.flat_grad.abs().max()) == 0.0
    assert eng.ops.fallbacks == {}, eng.ops.fallbacks
    if arch in ("efficientnet_b0", "regnety_160"):
        assert fake.calls.get("se_gate_fwd", 0) > 0 and fake.calls.get("se_gate_bwd", 0) == fake.calls["se_gate_fwd"]
    if arch == "resnext50_32x4d":
        assert fake.calls.get("blockdiag_pack", 0) > 0 and fake.calls.get("blockdiag_unpack_add", 0) > 0
    if arch not in ("densenet121", "efficientnet_b0"):   # strided dense/grouped convs: parity-class dgrad, no zero insertion
        assert fake.calls.get("conv_dgrad_s2", 0) + fake.calls.get("conv_dgrad_s2+addend", 0) + fake.calls.get("strided_add_inplace", 0) > 0


def test_activation_checkpointing_recomputes_on_the_native_path(cpu_engine):
    """DenseNet ``memory_efficient=True`` (reference densenet.py:82-86): the recomputation happens during backward,
    outside the engine's forward scope (and on autograd's worker thread on CUDA); it must re-enter the native path,
    reuse the BN statistic slots from zero and still match the torch path, which also recomputes."""
    eng, ref, fake = cpu_engine("densenet121", num_classes=16, memory_efficient=True)
    losses, _ = _train(eng, ref, steps=2, batch=8, size=64, classes=16, lr=0.005)
    for la, lb in losses:
        assert abs(la - lb) / max(abs(lb), 1e-3) < 0.05, losses
    sd_a, sd_b = eng.module.state_dict(), ref.state_dict()
    k = "features.denseblock1.denselayer1.norm1.num_batches_tracked"      # stepped by forward AND by the recomputation
    assert int(sd_a[k]) == int(sd_b[k]) == 4


def test_frozen_batchnorm_in_a_training_step(cpu_engine):
    """BN modules left in eval mode while training (fine-tuning with frozen statistics): running statistics are
    used and stay untouched, gradients flow without the batch-mean terms."""
    eng, ref, _ = cpu_engine("resnet18", num_classes=16)
    eng.train(), ref.train()
    for net in (eng.module, ref):
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
                m.running_var.fill_(4.0)
    before = eng.module.bn1.running_mean.clone()
    opt = eng.make_optimizer(lr=0.01, momentum=0.9, dampening=0.0, weight_decay=0.0, nesterov=True)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9, nesterov=True)
    init = {k: v.clone() for k, v in ref.state_dict().items()}
    x, y = torch.randn(8, 3, 64, 64), torch.randint(0, 16, (8,))
    la, _, _ = eng.train_step(x, y, opt, 5)
    lb, _, _ = Fn.cross_entropy_topk(ref(x), y, 5)
    ropt.zero_grad()
    lb.backward()
    ropt.step()
    assert abs(float(la.detach()) - float(lb.detach())) / float(lb.detach()) < 0.03
    assert torch.equal(eng.module.bn1.running_mean, before) and int(eng.module.bn1.num_batches_tracked) == 0
    sa, sb = eng.module.state_dict(), ref.state_dict()
    for k in ("conv1.weight", "layer2.0.downsample.0.weight", "layer4.1.bn2.weight", "fc.weight"):
        da, db = (sa[k].float() - init[k]).flatten(), (sb[k] - init[k]).flatten()
        assert float(torch.dot(da, db) / (da.norm() * db.norm())) > 0.9, k


def test_uint8_batches_match_host_normalisation(cpu_engine):
    from distribuuuu_b200.utils.data import normalize_uint8
    eng, _, fake = cpu_engine("resnet18", num_classes=16)
    eng.eval()
    x = torch.randint(0, 256, (2, 3, 64, 64), dtype=torch.uint8)
    y = torch.zeros(2, dtype=torch.long)
    with torch.no_grad():
        la, _, _ = eng.eval_step(x, y, 5)
        lb, _, _ = eng.eval_step(normalize_uint8(x), y, 5)
    assert abs(float(la) - float(lb)) < 1e-2 and fake.calls["stem_s2d"] == 2     # 7x7/2 stem: space-to-depth path
    # the explicit-im2col stem (what 3x3/2 stems use, and the fallback of the 7x7 one) gives the same answer
    eng.stem_s2d = False
    with torch.no_grad():
        lc, _, _ = eng.eval_step(x, y, 5)
    assert abs(float(la) - float(lc)) < 1e-2 and fake.calls["stem_im2col"] == 1


def test_eval_matches_torch_in_eval_mode(cpu_engine):
    eng, ref, _ = cpu_engine("resnet18", num_classes=16)
    eng.eval(), ref.eval()
    x = torch.randn(4, 3, 64, 64)
    y = torch.randint(0, 16, (4,))
    with torch.no_grad():
        la, _, _ = eng.eval_step(x, y, 5)
        lb, _, _ = Fn.cross_entropy_topk(ref(x), y, 5)
    assert abs(float(la) - float(lb)) / abs(float(lb)) < 0.03


def test_checkpoints_interchange_with_plain_torch(cpu_engine, fresh_cfg, tmp_path):
    """utils.save_checkpoint / load_checkpoint with the native engine and its fused optimizer: reference layout
    (utils.py:366-410), torch.optim.SGD-format optimizer state, and a lossless round trip through a plain model."""
    from distribuuuu_b200 import utils
    fresh_cfg.OUT_DIR = str(tmp_path)
    eng, _, _ = cpu_engine("resnet18", num_classes=16)
    opt = eng.make_optimizer(lr=0.05, momentum=0.9, dampening=0.0, weight_decay=5e-5, nesterov=True)
    eng.train()
    x, y = torch.randn(4, 3, 64, 64), torch.randint(0, 16, (4,))
    for _ in range(2):
        eng.train_step(x, y, opt, 5)
    path = utils.save_checkpoint(eng, opt, 0, 1.0, True)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"epoch", "state_dict", "optimizer", "best_acc1"}
    assert not any(k.startswith("module.") for k in ck["state_dict"])
    plain = build_model("resnet18", num_classes=16)
    popt = torch.optim.SGD(plain.parameters(), lr=0.1, momentum=0.9, nesterov=True)
    assert utils.load_checkpoint(path, plain, popt) == (1, 1.0)
    first = next(iter(plain.parameters()))
    assert popt.state[first]["momentum_buffer"].shape == first.shape
    # ... and back into a fresh engine: masters, momentum and bf16 compute weights are restored exactly
    eng2, _, _ = cpu_engine("resnet18", num_classes=16)
    opt2 = eng2.make_optimizer(lr=0.05, momentum=0.9, dampening=0.0, weight_decay=5e-5, nesterov=True)
    torch.save({"epoch": 0, "state_dict": plain.state_dict(), "optimizer": popt.state_dict(), "best_acc1": 1.0}, path)
    utils.load_checkpoint(path, eng2, opt2)
    assert torch.equal(eng2.flat_master, eng.flat_master) and torch.equal(eng2.flat_mom, eng.flat_mom)
    assert torch.equal(eng2.flat_w16, eng.flat_w16)
    # the restored engine continues exactly like the original
    la, _, _ = eng.train_step(x, y, opt, 5)
    lb, _, _ = eng2.train_step(x, y, opt2, 5)
    assert float(la.detach()) == float(lb.detach()) and torch.equal(eng.flat_master, eng2.flat_master)


def test_stem_paths_and_their_fallbacks(cpu_engine):
    """The 7x7/2 stem: space-to-depth + fused BN/ReLU/max-pool tail when the geometry allows; a conv output with an odd
    side takes the three separate ops; a kernel module / driver that rejects the space-to-depth map (RuntimeError) makes the
    engine fall back to the explicit-im2col stem for good, with the same loss."""
    eng, ref, fake = cpu_engine("resnet18", num_classes=16)
    opt = eng.make_optimizer(lr=0.01, momentum=0.9, dampening=0.0, weight_decay=0.0, nesterov=True)
    eng.train(), ref.train()
    g = torch.Generator().manual_seed(3)
    y = torch.randint(0, 16, (4,), generator=g)
    # 62x62 input -> 31x31 conv output: the pooling windows do not tile 2x2 blocks, separate BN / pool kernels
    x_odd = torch.randn(4, 3, 62, 62, generator=g)
    lr_, _, _ = Fn.cross_entropy_topk(ref(x_odd), y, 5)          # same (initial) weights as the engine
    lb, _, _ = eng.train_step(x_odd, y, opt, 5)
    assert fake.calls.get("stem_s2d", 0) == 1 and fake.calls.get("maxpool_fwd", 0) == 1 and fake.calls.get("maxpool_bwd", 0) == 1
    assert fake.calls.get("bn_relu_pool_fwd", 0) == 0
    assert abs(float(lb.detach()) - float(lr_.detach())) / float(lr_.detach()) < 0.05
    # 64x64 input -> 32x32 conv output: fused tail
    x = torch.randn(4, 3, 64, 64, generator=g)
    la, _, _ = eng.train_step(x, y, opt, 5)
    assert fake.calls.get("stem_s2d", 0) == 2 and fake.calls.get("bn_relu_pool_fwd", 0) == 1 and fake.calls.get("maxpool_fwd", 0) == 1
    assert fake.calls.get("bn_relu_pool_bwd", 0) == 1 and float(la.detach()) == float(la.detach())
    # the space-to-depth kernel refuses: explicit im2col from then on
    real = fake.stem_s2d

    def refuse(*a, **k):
        raise RuntimeError("cuTensorMapEncodeIm2col failed (emulated)")
    fake.stem_s2d = refuse
    before = fake.calls.get("stem_im2col", 0)
    lc, _, _ = eng.train_step(x, y, opt, 5)
    fake.stem_s2d = real
    assert eng.stem_s2d is False and fake.calls.get("stem_im2col", 0) == before + 1
    ld, _, _ = eng.train_step(x, y, opt, 5)
    assert fake.calls.get("stem_im2col", 0) == before + 2 and float(ld.detach()) == float(ld.detach())
