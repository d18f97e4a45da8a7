"""BoTNet-50: ResNet-50 whose last stage is three bottleneck-transformer blocks.

Parity: reference ``distribuuuu/models/botnet.py`` -- ``MHSA`` 163-215 (1x1 convs to q,k
(4 heads x 128) and v, scaled dot-product + 2-D relative position logits, softmax, PV),
``RelPosEmb``/``AbsPosEmb`` 60-98, ``BoTBlock`` 101-160, ``BoTStack`` 218-272 and the
assembly ``botnet50`` 275-290.  The module tree reproduces the reference's
``nn.Sequential`` indexing, so state_dict keys are identical
(``7.net.0.net.3.pos_emb.rel_height`` ...).

Differences by design: the relative-position logits are computed from index tables
instead of the pad/reshape trick with hard-coded ``.cuda()`` (reference 33,36), so the
model runs on any device; ``num_classes`` is honoured (the reference hard-codes 1000,
botnet.py:288); on the native path the attention core (QK^T + relative logits + softmax + PV, forward and backward)
runs as fused tcgen05 kernels on the NHWC projections (``Fn.relpos_mhsa`` -> ``csrc/attention.cu``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..ops import functional as Fn
from .resnet import resnet50

__all__ = ["botnet50", "BoTStack", "BoTBlock", "MHSA", "RelPosEmb", "AbsPosEmb"]


class RelPosEmb(nn.Module):
    def __init__(self, height: int, width: int, dim_head: int):
        super().__init__()
        self.height, self.width = height, width
        scale = dim_head ** -0.5
        self.rel_height = nn.Parameter(torch.randn(height * 2 - 1, dim_head) * scale)
        self.rel_width = nn.Parameter(torch.randn(width * 2 - 1, dim_head) * scale)


class AbsPosEmb(nn.Module):
    def __init__(self, height: int, width: int, dim_head: int):
        super().__init__()
        scale = dim_head ** -0.5
        self.height = nn.Parameter(torch.randn(height, dim_head) * scale)
        self.width = nn.Parameter(torch.randn(width, dim_head) * scale)


class MHSA(nn.Module):
    """All-to-all self-attention over an HxW feature map."""

    def __init__(self, dim, fmap_size, heads=4, dim_qk=128, dim_v=128, rel_pos_emb=False):
        super().__init__()
        self.scale = dim_qk ** -0.5
        self.heads, self.dim_qk, self.dim_v = heads, dim_qk, dim_v
        self.fmap_size = tuple(fmap_size)
        self.to_qk = nn.Conv2d(dim, heads * dim_qk * 2, 1, bias=False)
        self.to_v = nn.Conv2d(dim, heads * dim_v, 1, bias=False)
        self.softmax = nn.Softmax(dim=-1)
        h, w = self.fmap_size
        self.rel = bool(rel_pos_emb)
        self.pos_emb = RelPosEmb(h, w, dim_qk) if rel_pos_emb else AbsPosEmb(h, w, dim_qk)

    def forward(self, x):
        B, _, H, W = x.shape
        nh = self.heads
        qk = Fn.conv2d(x, self.to_qk)
        v = Fn.conv2d(x, self.to_v)
        if self.rel:   # fused on the native path (one tcgen05 kernel on the NHWC projections)
            return Fn.relpos_mhsa(qk, v, self.pos_emb.rel_height, self.pos_emb.rel_width, nh, self.dim_qk, self.dim_v, self.scale)
        q, k = qk[:, : nh * self.dim_qk], qk[:, nh * self.dim_qk:]

        def split(t, d):  # [B,(h d),H,W] -> [B,h,HW,d]
            return t.reshape(B, nh, d, H * W).transpose(2, 3)

        q, k, v = split(q, self.dim_qk), split(k, self.dim_qk), split(v, self.dim_v)
        out = Fn.abspos_attention(q, k, v, self.pos_emb.height, self.pos_emb.width, self.scale)
        return out.transpose(2, 3).reshape(B, nh * self.dim_v, H, W)


class BoTBlock(nn.Module):
    def __init__(self, dim, fmap_size, dim_out, stride=1, heads=4, proj_factor=4, dim_qk=128,
                 dim_v=128, rel_pos_emb=False, activation=None):
        super().__init__()
        act = activation if activation is not None else nn.ReLU()
        if dim != dim_out or stride != 1:
            self.shortcut = nn.Sequential(nn.Conv2d(dim, dim_out, 1, stride=stride, bias=False),
                                          nn.BatchNorm2d(dim_out), act)
        else:
            self.shortcut = nn.Identity()
        mid = dim_out // proj_factor
        attn_out = heads * dim_v
        self.stride = stride
        self.net = nn.Sequential(
            nn.Conv2d(dim, mid, 1, bias=False),                                    # 0
            nn.BatchNorm2d(mid),                                                   # 1
            act,                                                                   # 2
            MHSA(mid, fmap_size, heads, dim_qk, dim_v, rel_pos_emb),               # 3
            nn.AvgPool2d((2, 2)) if stride == 2 else nn.Identity(),                # 4
            nn.BatchNorm2d(attn_out),                                              # 5
            act,                                                                   # 6
            nn.Conv2d(attn_out, dim_out, 1, bias=False),                           # 7
            nn.BatchNorm2d(dim_out),                                               # 8
        )
        nn.init.zeros_(self.net[8].weight)  # zero-gamma on the block's last BN
        self.activation = act

    def forward(self, x):
        if isinstance(self.shortcut, nn.Identity):
            sc = x
        else:
            sc = Fn.conv_bn_act(x, self.shortcut[0], self.shortcut[1], "relu")
        n = self.net
        y = Fn.conv_bn_act(x, n[0], n[1], "relu")
        y = n[3](y)
        if self.stride == 2:
            y = Fn.avg_pool2d(y, 2, 2)
        y = Fn.bn_act(y, n[5], "relu")
        return Fn.conv_bn_act(y, n[7], n[8], "relu", residual=sc)


class BoTStack(nn.Module):
    def __init__(self, dim, fmap_size, dim_out=2048, heads=4, proj_factor=4, num_layers=3, stride=2,
                 dim_qk=128, dim_v=128, rel_pos_emb=False, activation=None):
        super().__init__()
        self.dim, self.fmap_size = dim, tuple(fmap_size)
        blocks = []
        for i in range(num_layers):
            first = i == 0
            div = 2 if (stride == 2 and not first) else 1
            blocks.append(BoTBlock(dim if first else dim_out, tuple(s // div for s in fmap_size), dim_out,
                                   stride if first else 1, heads, proj_factor, dim_qk, dim_v, rel_pos_emb,
                                   activation))
        self.net = nn.Sequential(*blocks)

    def forward(self, x):
        _, c, h, w = x.shape
        assert c == self.dim, f"channels of feature map {c} != BoTStack dim {self.dim}"
        assert (h, w) == self.fmap_size, f"feature map {(h, w)} != configured {self.fmap_size}"
        return self.net(x)


class BoTNet(nn.Sequential):
    """Index layout: 0 conv1, 1 bn1, 2 relu, 3 maxpool, 4-6 layer1-3, 7 BoTStack, 8 avgpool,
    9 flatten, 10 fc -- the same positions the reference's ``nn.Sequential`` produces."""

    def forward(self, x):
        x = Fn.conv_bn_relu_maxpool(x, self[0], self[1], 3, 2, 1)
        x = self[4](x)
        x = self[5](x)
        x = self[6](x)
        x = self[7](x)
        return Fn.linear(Fn.global_avg_pool(x), self[10])


def botnet50(pretrained=False, num_classes=1000, fmap_size=(14, 14), **kwargs):
    """BoTNet-50 for 224x224 inputs (``fmap_size`` = input / 16)."""
    trunk = resnet50(pretrained=pretrained, **kwargs)
    stack = BoTStack(dim=1024, fmap_size=fmap_size, stride=1, rel_pos_emb=True)
    return BoTNet(trunk.conv1, trunk.bn1, trunk.relu, trunk.maxpool, trunk.layer1, trunk.layer2, trunk.layer3,
                  stack, nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten(1), nn.Linear(2048, num_classes))


def smoke_botnet50(device="cpu"):
    """Shape check of the full model (reference ``test_botnet50``, botnet.py:293-297; device-agnostic here)."""
    import torch
    x = torch.ones(2, 3, 224, 224, device=device)
    y = botnet50().to(device)(x)
    assert tuple(y.shape) == (2, 1000)
    return tuple(y.shape)


def smoke_backbone(device="cpu"):
    """BoTStack dropped into a ResNet-50 trunk keeps the backbone's output shape (reference ``test_backbone``,
    botnet.py:300-314)."""
    import torch
    trunk = resnet50()
    layers = [trunk.conv1, trunk.bn1, trunk.relu, trunk.maxpool, trunk.layer1, trunk.layer2, trunk.layer3,
              BoTStack(dim=1024, fmap_size=(14, 14), stride=1, rel_pos_emb=True)]
    x = torch.ones(2, 3, 224, 224, device=device)
    net = nn.Sequential(*layers).to(device)
    y = x
    for m in net:
        y = m(y)
    assert tuple(y.shape) == (2, 2048, 14, 14)
    return tuple(y.shape)


if __name__ == "__main__":
    print(smoke_backbone())
