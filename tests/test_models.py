"""Model zoo: parameter counts (reference README.md:206-217), key naming, forward shapes."""
import pytest
import torch

from distribuuuu_b200 import models
from distribuuuu_b200.ops import functional as Fn

PARAMS_M = dict(resnet18=(11.690, 62), resnet34=(21.798, 110), resnet50=(25.557, 161), resnet101=(44.549, 314),
                resnet152=(60.193, 467), resnext50_32x4d=(25.029, 161), resnext101_32x8d=(88.791, 314),
                wide_resnet50_2=(68.883, 161), wide_resnet101_2=(126.887, 314), densenet121=(7.979, 364),
                densenet169=(14.149, 508), densenet201=(20.014, 604), densenet161=(28.681, 484),
                botnet50=(20.859, 170), regnetx_160=(54.279, 215), regnety_160=(83.590, 251),
                regnety_320=(145.047, 277), efficientnet_b0=(5.289, 213))


@pytest.mark.parametrize("arch", sorted(PARAMS_M))
def test_param_counts(arch):
    m = models.build_model(arch)
    n = sum(p.numel() for p in m.parameters())
    assert round(n / 1e6, 3) == PARAMS_M[arch][0]
    assert len(list(m.parameters())) == PARAMS_M[arch][1]


def test_registry_contract():
    assert set(PARAMS_M) <= set(models.list_models())
    with pytest.raises(KeyError):
        models.build_model("not_a_model")
    assert models.build_model("resnet18", num_classes=7).fc.out_features == 7
    with pytest.raises(RuntimeError):
        models.build_model("efficientnet_b0", pretrained=True)


@pytest.mark.parametrize("arch,size", [("resnet18", 64), ("resnet50", 64), ("resnext50_32x4d", 64),
                                       ("densenet121", 64), ("regnetx_160", 64), ("regnety_160", 64),
                                       ("efficientnet_b0", 64), ("botnet50", 224)])
def test_forward_backward_shapes(arch, size):
    torch.manual_seed(0)
    m = models.build_model(arch, num_classes=11)
    out = m(torch.randn(2, 3, size, size))
    assert out.shape == (2, 11)
    out.sum().backward()
    missing = [n for n, p in m.named_parameters() if p.grad is None]
    assert not missing, missing[:5]


def test_botnet_keys_match_reference_sequential_layout():
    keys = set(models.build_model("botnet50").state_dict())
    for k in ["0.weight", "1.running_mean", "4.0.conv1.weight", "6.5.bn3.weight", "7.net.0.shortcut.0.weight",
              "7.net.0.net.3.to_qk.weight", "7.net.0.net.3.pos_emb.rel_height", "7.net.2.net.3.pos_emb.rel_width",
              "7.net.1.net.8.weight", "10.weight", "10.bias"]:
        assert k in keys, k
    assert not any(k.startswith("7.net.1.shortcut") for k in keys)  # identity shortcut has no params


def test_timm_style_keys_for_regnet_and_efficientnet():
    rk = set(models.build_model("regnety_160").state_dict())
    for k in ["stem.conv.weight", "stem.bn.weight", "s1.b1.conv1.conv.weight", "s1.b1.conv2.bn.running_var",
              "s1.b1.se.fc1.bias", "s1.b1.downsample.conv.weight", "s4.b1.conv3.bn.weight", "head.fc.weight"]:
        assert k in rk, k
    ek = set(models.build_model("efficientnet_b0").state_dict())
    for k in ["conv_stem.weight", "bn1.weight", "blocks.0.0.conv_dw.weight", "blocks.0.0.se.conv_reduce.bias",
              "blocks.1.0.conv_pw.weight", "blocks.6.0.conv_pwl.weight", "blocks.6.0.bn3.bias", "conv_head.weight",
              "bn2.running_mean", "classifier.bias"]:
        assert k in ek, k


def test_regnety_se_widths_follow_block_input():
    m = models.build_model("regnety_160")
    assert m.s1.b1.se.fc1.out_channels == 8 and m.s1.b2.se.fc1.out_channels == 56
    assert m.s3.b1.se.fc1.out_channels == 112 and m.s3.b2.se.fc1.out_channels == 308
    assert m.s2.b1.conv2.conv.groups == 4 and m.s4.b1.conv2.conv.groups == 27


def _relpos_bruteforce(q, k, v, rel_h, rel_w, H, W, scale):
    B, nh, L, d = q.shape
    q = q * scale
    logits = q @ k.transpose(-1, -2)
    for x in range(H):
        for y in range(W):
            for i in range(H):
                for j in range(W):
                    r = rel_w[j - y + W - 1] + rel_h[i - x + H - 1]
                    logits[:, :, x * W + y, i * W + j] += (q[:, :, x * W + y] * r).sum(-1)
    return torch.softmax(logits, -1) @ v


def test_relpos_attention_matches_bruteforce_and_padreshape_trick():
    torch.manual_seed(1)
    B, nh, H, W, d = 2, 2, 3, 4, 8
    q, k, v = (torch.randn(B, nh, H * W, d) for _ in range(3))
    rel_h, rel_w = torch.randn(2 * H - 1, d), torch.randn(2 * W - 1, d)
    got = Fn.relpos_attention(q, k, v, rel_h, rel_w, H, W, d ** -0.5)
    want = _relpos_bruteforce(q.clone(), k, v, rel_h, rel_w, H, W, d ** -0.5)
    assert torch.allclose(got, want, atol=1e-5)

    # the pad / flatten / reshape construction of relative->absolute indexing (Bello et al. 2019,
    # used by the reference botnet.py:25-40) gives the same width logits as the index table
    L = W
    rel_logits = torch.einsum("bnxyd,md->bnxym", (q * d ** -0.5).reshape(B, nh, H, W, d), rel_w)
    x = rel_logits.reshape(B, nh * H, L, 2 * L - 1)
    x = torch.cat([x, x.new_zeros(B, nh * H, L, 1)], 3).reshape(B, nh * H, L * 2 * L)
    x = torch.cat([x, x.new_zeros(B, nh * H, L - 1)], 2).reshape(B, nh * H, L + 1, 2 * L - 1)[:, :, :L, L - 1:]
    iw = torch.arange(W)
    tab = rel_w[(iw[None, :] - iw[:, None]) + W - 1]
    mine = torch.einsum("bnxyd,yjd->bnxyj", (q * d ** -0.5).reshape(B, nh, H, W, d), tab)
    assert torch.allclose(x.reshape(B, nh, H, W, W), mine, atol=1e-5)


def test_densenet_memory_efficient_matches():
    torch.manual_seed(0)
    a = models.build_model("densenet121", num_classes=5)
    b = models.build_model("densenet121", num_classes=5, memory_efficient=True)
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 3, 64, 64, requires_grad=True)
    ya, yb = a(x), b(x)
    assert torch.allclose(ya, yb, atol=1e-5)
    yb.sum().backward()
    assert b.features.denseblock1.denselayer1.conv1.weight.grad is not None
