"""GPU numerics: every sm_100a kernel against a plain PyTorch fp32 reference of the same op (run on the B200 box)."""
import pytest

pytestmark = pytest.mark.gpu


def test_extension_loads_and_is_in_tree():
    import os
    from distribuuuu_b200.ops import build
    mod = build.load()
    assert os.path.dirname(os.path.abspath(mod.__file__)).endswith(os.path.join("distribuuuu_b200", "_ext"))


def _cases():
    from distribuuuu_b200.selftest import CONV_CASES
    return sorted(CONV_CASES)


@pytest.mark.parametrize("name", _cases())
def test_conv_gemm_tcgen05(name):
    from distribuuuu_b200 import selftest
    selftest.check_conv_case(name)


@pytest.mark.parametrize("act,residual,C", [("relu", True, 64), (None, False, 256), ("silu", False, 24), ("relu", False, 2048)])
def test_batchnorm_kernels(act, residual, C):
    from distribuuuu_b200 import selftest
    selftest.check_bn(act=act, residual=residual, C=C)


def test_batchnorm_relu_bitmask_backward():
    from distribuuuu_b200 import selftest
    selftest.check_bn(act="relu", residual=True, C=256, use_mask=True)


@pytest.mark.parametrize("C,G,stride", [(224, 2, 1), (512, 4, 2), (696, 3, 1)])
def test_grouped_conv_tcgen05(C, G, stride):
    from distribuuuu_b200 import selftest
    selftest.check_grouped_conv(C=C, K=C, G=G, stride=stride, H=28 if stride == 2 else 14, W=28 if stride == 2 else 14)


@pytest.mark.parametrize("k,stride,C", [(3, 1, 32), (5, 2, 96), (3, 2, 144)])
def test_depthwise_conv_kernels(k, stride, C):
    from distribuuuu_b200 import selftest
    selftest.check_depthwise(k=k, stride=stride, C=C)


def test_squeeze_excite_scale_kernels():
    from distribuuuu_b200 import selftest
    selftest.check_channel_scale()
    selftest.check_channel_scale(N=2, H=14, W=14, C=1232)


def test_pool_kernels():
    from distribuuuu_b200 import selftest
    selftest.check_pools()


@pytest.mark.parametrize("kw", [dict(), dict(N=3, H=10, W=14, C=96), dict(N=64, H=112, W=112, C=64), dict(N=2, H=4, W=600, C=64),
                                dict(N=2, H=6, W=10, C=24)],
                         ids=["small", "c96_odd", "stem_shape", "wide_rows_register_fallback", "c24_unaligned_arg_rows"])
def test_fused_stem_tail_bn_relu_maxpool(kw):
    from distribuuuu_b200 import selftest
    selftest.check_bn_relu_pool(**kw)


def test_softmax_ce_topk_kernel():
    from distribuuuu_b200 import selftest
    selftest.check_ce_topk()


def test_fused_sgd_matches_torch_optim():
    from distribuuuu_b200 import selftest
    selftest.check_sgd()


def test_engine_accepts_uint8_batches():
    from distribuuuu_b200 import selftest
    selftest.check_uint8_input()


def test_stem_im2col_and_layout_conversion():
    from distribuuuu_b200 import selftest
    selftest.check_stem()


@pytest.mark.parametrize("kw", [dict(), dict(N=16, H=224, W=224), dict(N=3, H=96, W=160, Kc=96)], ids=["64", "224", "96x160_k96"])
def test_stem_space_to_depth_on_im2col_kernels(kw):
    from distribuuuu_b200 import selftest
    selftest.check_stem_s2d(**kw)


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_native_engine_tracks_fp32_reference(arch):
    from distribuuuu_b200 import selftest
    # 16 x 96^2 and a small step: with 8 x 64^2 the last BN layers see 32 samples per channel and lr 0.01 blows the loss
    # up to ~10 within three steps -- a chaotic regime in which the run-to-run noise of the fp32 atomics alone moved the
    # deviation between 1 % and 8 % (tools/rep_engine_check.py), whatever stem path was used
    selftest.check_engine_vs_torch(arch, batch=16, size=96, lr=0.002)


@pytest.mark.parametrize("arch,batch,size", [("resnext50_32x4d", 8, 64), ("densenet121", 8, 64), ("efficientnet_b0", 16, 128),
                                             ("regnety_160", 4, 64), ("regnetx_160", 4, 64), ("botnet50", 4, 224)])
def test_every_model_family_trains_on_native_engine(arch, batch, size):
    from distribuuuu_b200 import selftest
    # EfficientNet's SiLU/SE stack is chaotic at random init with a large step; a small LR keeps the 3-step
    # trajectories of the bf16 and fp32 paths comparable
    selftest.check_engine_vs_torch(arch, batch=batch, size=size, tol=0.15, lr=0.005 if arch == "efficientnet_b0" else 0.01)


@pytest.mark.parametrize("arch,batch,size", [("resnet50", 16, 128), ("efficientnet_b0", 16, 128), ("regnety_160", 8, 128)])
def test_native_gradients_match_fp32_per_parameter(arch, batch, size):
    from distribuuuu_b200 import selftest
    selftest.check_engine_grads(arch, batch=batch, size=size)


def test_checkpoint_interop_native_vs_torch_optim():
    from distribuuuu_b200 import selftest
    selftest.check_checkpoint_interop()


def test_smoke_entry():
    import __graft_entry__
    __graft_entry__.smoke()


@pytest.mark.parametrize("kw", [dict(), dict(N=3, H=7, W=7, C=64, K=192, with_addend=True), dict(N=2, H=28, W=28, C=256, K=512, R=1, pad=0),
                                dict(N=2, H=14, W=14, C=64, K=128, R=1, pad=0, with_addend=True), dict(N=2, H=28, W=28, C=224, K=224, G=2),
                                dict(N=2, H=12, W=12, C=64, K=64, R=5, pad=2)],
                         ids=["3x3", "3x3_odd_addend", "1x1", "1x1_addend", "grouped", "5x5"])
def test_stride2_dgrad_by_parity_classes(kw):
    from distribuuuu_b200 import selftest
    selftest.check_dgrad_s2(**kw)


@pytest.mark.parametrize("kw", [dict(), dict(C=256, G=32, H=28, W=28, stride=2), dict(C=512, G=32, H=7, W=7)], ids=["cg4", "cg8_s2", "cg16"])
def test_thin_group_convs_block_diagonal(kw):
    from distribuuuu_b200 import selftest
    selftest.check_thin_groups(**kw)


@pytest.mark.parametrize("kw", [dict(), dict(N=9, H=4, W=4, C=1232, r=308, act="relu"), dict(N=33, H=7, W=7, C=480, r=20)],
                         ids=["silu_r4", "relu_r308", "silu_r20"])
def test_fused_squeeze_excite(kw):
    from distribuuuu_b200 import selftest
    selftest.check_se(**kw)


def test_bias_gradient_colsum():
    from distribuuuu_b200 import selftest
    selftest.check_colsum()


@pytest.mark.parametrize("arch", ["resnet50", "resnext50_32x4d", "regnety_160", "efficientnet_b0"])
def test_no_library_fallbacks_for_zoo_models(arch):
    """One native training step of a config/*.yaml family must never leave the sm_100a kernels."""
    import torch
    from distribuuuu_b200 import models
    from distribuuuu_b200.parallel.native_engine import NativeEngine
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = models.build_model(arch, num_classes=16).to(dev)
    eng = NativeEngine(net, dev)
    opt = eng.make_optimizer(lr=0.01, momentum=0.9, dampening=0.0, weight_decay=5e-5, nesterov=True)
    eng.train()
    x = torch.randn(4, 3, 64, 64, device=dev)
    y = torch.randint(0, 16, (4,), device=dev)
    loss, _, _ = eng.train_step(x, y, opt, 5)
    assert torch.isfinite(loss).item()
    assert eng.ops.fallbacks == {}, eng.ops.fallbacks


@pytest.mark.parametrize("B", [1, 3])
def test_fused_relpos_mhsa(B):
    from distribuuuu_b200 import selftest
    selftest.check_mhsa(B=B)


def test_botnet50_step_has_no_library_fallbacks():
    import torch
    from distribuuuu_b200 import models
    from distribuuuu_b200.parallel.native_engine import NativeEngine
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = models.build_model("botnet50", num_classes=16).to(dev)
    eng = NativeEngine(net, dev)
    opt = eng.make_optimizer(lr=0.01, momentum=0.9, dampening=0.0, weight_decay=5e-5, nesterov=True)
    eng.train()
    x = torch.randn(2, 3, 224, 224, device=dev)
    y = torch.randint(0, 16, (2,), device=dev)
    loss, _, _ = eng.train_step(x, y, opt, 5)
    assert torch.isfinite(loss).item()
    assert eng.ops.fallbacks == {}, eng.ops.fallbacks


@pytest.mark.parametrize("N,H,W", [(2, 56, 56), (3, 14, 14), (5, 9, 13), (256, 56, 56)])
def test_conv3x3_halo_kernel(N, H, W):
    from distribuuuu_b200 import selftest
    selftest.check_conv_halo(N=N, H=H, W=W)
