"""Host-side logic of the native engine that needs no GPU: bucket planning over the flat layout and the
torch.optim.SGD-compatible optimizer state of FusedSGD (reference checkpoint layout, utils.py:375-380)."""
import torch
import torch.nn as nn

from distribuuuu_b200.models import build_model
from distribuuuu_b200.parallel.native_engine import _ALIGN, FusedSGD, NativeEngine


class _HostEngine(NativeEngine):
    """NativeEngine with the CUDA-dependent construction replaced by CPU tensors (layout logic only)."""

    def __init__(self, module):
        nn.Module.__init__(self)
        self.module = module
        self.world, self.rank, self.comm_mode = 1, 0, "local"
        self.device = torch.device("cpu")
        self._build_flat_storage()
        self.flat_w16 = torch.zeros(self.total, dtype=torch.bfloat16)

    def sync_masters(self):
        pass


def test_flat_layout_keeps_state_dict_and_alignment():
    net = build_model("resnet18", num_classes=10)
    want = {k: v.clone() for k, v in net.state_dict().items()}
    eng = _HostEngine(net)
    got = net.state_dict()
    assert all(torch.equal(got[k], want[k]) for k in want)              # values survive the re-binding
    for p in eng.params:
        off, n = eng.index[p]
        assert off % _ALIGN == 0 and n == p.numel()
        assert p.data.data_ptr() == eng.flat_master[off:].data_ptr()    # parameters are views of the flat master
    conv = net.conv1.weight                                             # 4-D weights are physically [O,H,W,I]
    off, n = eng.index[conv]
    O, I, H, W = conv.shape
    assert torch.equal(eng.flat_master[off:off + n].view(O, H, W, I), conv.data.permute(0, 2, 3, 1))


def test_buckets_cover_all_parameters_in_reverse_order():
    net = build_model("resnet50")
    eng = _HostEngine(net)
    eng._plan_buckets(25 * 1024 * 1024)
    assert (eng.buckets[0].n * 4) <= (1 << 20) + 4 * 2048 * 1000        # first bucket ~1 MiB (fc weight + ...)
    covered = sorted((b.off, b.off + b.n) for b in eng.buckets)
    assert covered[0][0] == 0 and covered[-1][1] == eng.trainable_total == eng.total
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))      # contiguous, non-overlapping
    assert eng.buckets[0].params[0] is net.fc.weight                     # gradients become ready last-layer first
    assert sum(len(b.params) for b in eng.buckets) == len(eng.params)
    assert all(b.n % 8 == 0 and b.off % 8 == 0 for b in eng.buckets)    # 16-byte vectors in the fused kernel


def test_one_dimensional_parameters_live_in_one_shot_buckets():
    """BN gamma/beta and biases are consumed in fp32 from the master buffer by the BN / bias kernels, and only a
    one-shot bucket keeps every rank's fp32 replica current (a two-shot bucket updates the owner's shard and
    broadcasts bf16) -- so no 1-D parameter may ever share a two-shot bucket."""
    for arch in ("resnet50", "regnety_160", "efficientnet_b0"):
        net = build_model(arch)
        eng = _HostEngine(net)
        eng._plan_buckets(25 * 1024 * 1024)
        for p in eng.params:
            b = eng.bucket_of[p]
            if p.dim() < 2:
                assert b.one_shot and b.n * 2 <= 2 * 512 * 1024 + 128, (arch, b.n)
                assert eng.index[p][0] >= eng.big_total
            else:
                assert eng.index[p][0] < eng.big_total
        assert any(not b.one_shot for b in eng.buckets)                  # the conv weights still go two-shot


def test_frozen_parameters_are_outside_the_fused_update_range():
    net = build_model("resnet18", num_classes=10)
    for p in net.layer1.parameters():
        p.requires_grad_(False)
    eng = _HostEngine(net)
    eng._plan_buckets(25 * 1024 * 1024)
    frozen = [p for p in eng.params if not p.requires_grad]
    assert frozen and all(eng.index[p][0] >= eng.trainable_total for p in frozen)
    assert all(p not in eng.bucket_of for p in frozen)
    assert max(b.off + b.n for b in eng.buckets) == eng.trainable_total < eng.total


def test_fused_sgd_state_dict_is_torch_sgd_compatible():
    net = build_model("resnet18", num_classes=10)
    eng = _HostEngine(net)
    opt = FusedSGD(eng, lr=0.1, momentum=0.9, dampening=0.0, weight_decay=5e-5, nesterov=True)
    eng.flat_mom.uniform_(-1, 1)
    opt.has_momentum_state = True
    sd = opt.state_dict()
    plain = build_model("resnet18", num_classes=10)
    topt = torch.optim.SGD(plain.parameters(), lr=0.5, momentum=0.9, nesterov=True)
    topt.load_state_dict(sd)                                            # the reference's optimizer accepts it
    assert topt.param_groups[0]["lr"] == 0.1 and topt.param_groups[0]["weight_decay"] == 5e-5
    for p_native, p_plain in zip(eng.params, plain.parameters()):
        buf = topt.state[p_plain]["momentum_buffer"]
        assert buf.shape == p_plain.shape
        assert torch.equal(buf, eng.logical_view(eng.flat_mom, p_native))
    # and the reverse direction: torch.optim.SGD state -> flat momentum
    eng2 = _HostEngine(build_model("resnet18", num_classes=10))
    opt2 = FusedSGD(eng2, lr=0.3, momentum=0.9, nesterov=True)
    opt2.load_state_dict(topt.state_dict())
    assert opt2.has_momentum_state and opt2.param_groups[0]["lr"] == 0.1
    for pa, pb in zip(eng.params, eng2.params):                          # (alignment padding between parameters is don't-care)
        assert torch.equal(eng.logical_view(eng.flat_mom, pa), eng2.logical_view(eng2.flat_mom, pb))


def test_grad_sink_mailbox_protocol():
    """ops.native._GradSink: one pending gradient at a time; late offers are refused so the producer falls back to
    returning its gradient through autograd (never a lost or double-counted contribution)."""
    from distribuuuu_b200.ops.native import _GradSink
    s = _GradSink(ptr=1234)
    assert s.take() is None and s.consumed            # consumer ran first ...
    assert not s.offer(torch.ones(1))                 # ... so a later producer must keep its gradient
    s = _GradSink(ptr=1234)
    g = torch.ones(2)
    assert s.offer(g) and not s.offer(torch.zeros(2))  # second producer is refused while one is pending
    assert s.take() is g and s.take() is None


def test_projection_shortcut_hint_is_ignored_on_the_torch_path():
    """models pass ``input_grad_to`` / ``residual_sink`` hints; without an engine they must not change results."""
    import copy
    from distribuuuu_b200.ops import functional as Fn
    torch.manual_seed(0)
    net = build_model("resnet18", num_classes=10)
    ref = copy.deepcopy(net)
    x = torch.randn(2, 3, 64, 64)
    blk, rblk = net.layer2[0], ref.layer2[0]          # block with a projection shortcut
    h = torch.randn(2, 64, 16, 16, requires_grad=True)
    h2 = h.detach().clone().requires_grad_(True)
    out = blk(h)
    identity = rblk.downsample[1](rblk.downsample[0](h2))
    want = torch.relu(rblk.bn2(rblk.conv2(torch.relu(rblk.bn1(rblk.conv1(h2))))) + identity)
    assert torch.allclose(out, want, atol=1e-5)
    out.sum().backward(); want.sum().backward()
    assert torch.allclose(h.grad, h2.grad, atol=1e-5)
    assert Fn.conv_bn_act(x, net.conv1, net.bn1, "relu", input_grad_to=net.conv1).shape == (2, 64, 32, 32)


def test_fused_stem_tail_emulation_matches_autograd():
    """BN -> ReLU -> max-pool fused op: the emulated kernels (and with them the check the GPU suite runs on the real
    ones) against F.batch_norm / relu / max_pool2d autograd."""
    from fake_kernels import FakeKernels
    from distribuuuu_b200 import selftest
    errs = selftest.check_bn_relu_pool(N=3, H=12, W=8, C=16, kmod=FakeKernels(), dev="cpu")
    assert errs["dy"] < 2e-2
