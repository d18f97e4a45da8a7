"""EfficientNet-B0.

Resolved through ``timm.create_model("efficientnet_b0", ...)`` in the reference
(trainer.py:124-128; config/efficientnet_b0.yaml).  Specified here from the published
architecture: stem 3x3/2 -> 32, seven MBConv stages
(expand, kernel, stride, out, repeats) = (1,3,1,16,1) (6,3,2,24,2) (6,5,2,40,2)
(6,3,2,80,3) (6,5,1,112,3) (6,5,2,192,4) (6,3,1,320,1), SE ratio 0.25 of the block input,
SiLU, head 1x1 -> 1280, no dropout / stochastic depth (timm defaults when the trainer
passes none).  Names follow timm (``conv_stem``, ``blocks.S.I.conv_pw`` ...,
``conv_head``, ``classifier``).  5.289 M parameters (reference README.md:212).
"""
from __future__ import annotations

import math

import torch.nn as nn

from ..ops import functional as Fn

__all__ = ["EfficientNet", "efficientnet_b0"]


class SqueezeExcite(nn.Module):
    def __init__(self, channels, rd_channels):
        super().__init__()
        self.conv_reduce = nn.Conv2d(channels, rd_channels, 1, bias=True)
        self.conv_expand = nn.Conv2d(rd_channels, channels, 1, bias=True)

    def forward(self, x):
        return Fn.squeeze_excite(x, self.conv_reduce, self.conv_expand, "silu")


class DepthwiseSeparable(nn.Module):
    """expand=1 MBConv: dw kxk -> SE -> 1x1 (linear)."""

    def __init__(self, cin, cout, k, stride, se_ratio=0.25):
        super().__init__()
        self.conv_dw = nn.Conv2d(cin, cin, k, stride=stride, padding=k // 2, groups=cin, bias=False)
        self.bn1 = nn.BatchNorm2d(cin)
        self.se = SqueezeExcite(cin, max(1, int(cin * se_ratio)))
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.has_skip = stride == 1 and cin == cout

    def forward(self, x):
        y = Fn.conv_bn_act(x, self.conv_dw, self.bn1, "silu")
        y = self.se(y)
        return Fn.conv_bn_act(y, self.conv_pw, self.bn2, None, residual=x if self.has_skip else None)


class InvertedResidual(nn.Module):
    """MBConv: 1x1 expand -> dw kxk -> SE -> 1x1 project (linear) [+ skip]."""

    def __init__(self, cin, cout, k, stride, expand, se_ratio=0.25):
        super().__init__()
        mid = cin * expand
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.conv_dw = nn.Conv2d(mid, mid, k, stride=stride, padding=k // 2, groups=mid, bias=False)
        self.bn2 = nn.BatchNorm2d(mid)
        self.se = SqueezeExcite(mid, max(1, int(cin * se_ratio)))
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.has_skip = stride == 1 and cin == cout

    def forward(self, x):
        y = Fn.conv_bn_act(x, self.conv_pw, self.bn1, "silu")
        y = Fn.conv_bn_act(y, self.conv_dw, self.bn2, "silu")
        y = self.se(y)
        return Fn.conv_bn_act(y, self.conv_pwl, self.bn3, None, residual=x if self.has_skip else None)


_B0_STAGES = ((1, 3, 1, 16, 1), (6, 3, 2, 24, 2), (6, 5, 2, 40, 2), (6, 3, 2, 80, 3),
              (6, 5, 1, 112, 3), (6, 5, 2, 192, 4), (6, 3, 1, 320, 1))


class EfficientNet(nn.Module):
    def __init__(self, stages=_B0_STAGES, stem_width=32, head_width=1280, num_classes=1000, drop_rate=0.0):
        super().__init__()
        self.conv_stem = nn.Conv2d(3, stem_width, 3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(stem_width)
        blocks, cin = [], stem_width
        for expand, k, stride, cout, reps in stages:
            stage = []
            for r in range(reps):
                s = stride if r == 0 else 1
                stage.append(DepthwiseSeparable(cin, cout, k, s) if expand == 1
                             else InvertedResidual(cin, cout, k, s, expand))
                cin = cout
            blocks.append(nn.Sequential(*stage))
        self.blocks = nn.Sequential(*blocks)
        self.conv_head = nn.Conv2d(cin, head_width, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(head_width)
        self.classifier = nn.Linear(head_width, num_classes)
        self.drop_rate = float(drop_rate)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan_out))
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                bound = 1.0 / math.sqrt(m.weight.size(0))
                nn.init.uniform_(m.weight, -bound, bound)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        x = Fn.conv_bn_act(x, self.conv_stem, self.bn1, "silu")
        x = self.blocks(x)
        x = Fn.conv_bn_act(x, self.conv_head, self.bn2, "silu")
        x = Fn.global_avg_pool(x)
        x = Fn.dropout(x, self.drop_rate, self.training)
        return Fn.linear(x, self.classifier)


def efficientnet_b0(pretrained=False, **kw):
    if pretrained:
        raise RuntimeError("efficientnet_b0: pretrained weights come from timm's hub in the reference; "
                           "no network here -- pass MODEL.WEIGHTS instead")
    return EfficientNet(**kw)
