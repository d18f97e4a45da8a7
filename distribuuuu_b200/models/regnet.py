"""RegNetX-16GF, RegNetY-16GF, RegNetY-32GF.

The reference does not define these: ``trainer.py:117-128`` falls back to
``timm.create_model`` for ``regnetx_160`` / ``regnety_160`` / ``regnety_320``
(config/regnet*.yaml).  timm is not available here, so the architectures are specified
from the design-space parameters: stem 3x3/2 -> 32; four stages, each
opening with a stride-2 block with a 1x1/2 projection shortcut; X block = 1x1 -> grouped
3x3 -> 1x1 (bottleneck ratio 1), Y block adds squeeze-excite after the 3x3 with squeeze
width = round(input_width / 4).  Module names follow timm (``stem.conv``,
``s1.b1.conv1.conv``, ``s1.b1.se.fc1``, ``head.fc``) so timm-format checkpoints map
onto them.  Parameter counts: 54.279 M / 83.590 M / 145.047 M (reference README.md:215-217).
"""
from __future__ import annotations

import math

import torch.nn as nn

from ..ops import functional as Fn

__all__ = ["RegNet", "regnetx_160", "regnety_160", "regnety_320"]


class ConvBnAct(nn.Module):
    def __init__(self, cin, cout, k=1, stride=1, groups=1, act="relu"):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, groups=groups, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        self.act = act

    def forward(self, x, residual=None, act="__default__", residual_sink=None, input_grad_to=None):
        return Fn.conv_bn_act(x, self.conv, self.bn, self.act if act == "__default__" else act, residual,
                              residual_sink=residual_sink, input_grad_to=input_grad_to)


class SEModule(nn.Module):
    def __init__(self, channels, rd_channels, act="relu"):
        super().__init__()
        self.fc1 = nn.Conv2d(channels, rd_channels, 1, bias=True)
        self.fc2 = nn.Conv2d(rd_channels, channels, 1, bias=True)
        self.act = act

    def forward(self, x):
        return Fn.squeeze_excite(x, self.fc1, self.fc2, self.act)


class RegBottleneck(nn.Module):
    def __init__(self, w_in, w_out, stride, group_width, se_ratio):
        super().__init__()
        groups = w_out // group_width
        self.conv1 = ConvBnAct(w_in, w_out, 1)
        self.conv2 = ConvBnAct(w_out, w_out, 3, stride=stride, groups=groups)
        self.se = SEModule(w_out, int(round(w_in * se_ratio))) if se_ratio else None
        self.conv3 = ConvBnAct(w_out, w_out, 1, act=None)
        self.downsample = ConvBnAct(w_in, w_out, 1, stride=stride, act=None) if (w_in != w_out or stride != 1) else None

    def forward(self, x):
        shortcut = x if self.downsample is None else self.downsample(x)
        # gradient-fusion hints (ops.functional.conv_bn_act): conv1 and the shortcut read the same tensor
        proj = self.downsample.conv if self.downsample is not None else None
        y = self.conv2(self.conv1(x, input_grad_to=proj))
        if self.se is not None:
            y = self.se(y)
        return self.conv3(y, residual=shortcut, act="relu", residual_sink=self.conv1.conv if proj is None else None)


class RegStage(nn.Sequential):
    def __init__(self, depth, w_in, w_out, group_width, se_ratio):
        super().__init__()
        for i in range(depth):
            self.add_module(f"b{i + 1}", RegBottleneck(w_in if i == 0 else w_out, w_out, 2 if i == 0 else 1,
                                                       group_width, se_ratio))


class _Head(nn.Module):
    def __init__(self, cin, num_classes):
        super().__init__()
        self.fc = nn.Linear(cin, num_classes)

    def forward(self, x):
        return Fn.linear(Fn.global_avg_pool(x), self.fc)


class RegNet(nn.Module):
    def __init__(self, depths, widths, group_width, se_ratio=0.0, num_classes=1000, stem_width=32,
                 zero_init_last_bn=True):
        super().__init__()
        self.stem = ConvBnAct(3, stem_width, 3, stride=2)
        w_prev = stem_width
        for i, (d, w) in enumerate(zip(depths, widths)):
            self.add_module(f"s{i + 1}", RegStage(d, w_prev, w, group_width, se_ratio))
            w_prev = w
        self.head = _Head(w_prev, num_classes)
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
                nn.init.normal_(m.weight, 0.0, 0.01)
                nn.init.zeros_(m.bias)
        if zero_init_last_bn:
            for m in self.modules():
                if isinstance(m, RegBottleneck):
                    nn.init.zeros_(m.conv3.bn.weight)

    def forward(self, x):
        x = self.stem(x)
        x = self.s1(x)
        x = self.s2(x)
        x = self.s3(x)
        x = self.s4(x)
        return self.head(x)


def _no_pretrained(name, pretrained):
    if pretrained:
        raise RuntimeError(f"{name}: pretrained weights come from timm's hub in the reference "
                           "(trainer.py:124-128); no network here -- pass MODEL.WEIGHTS instead")


def regnetx_160(pretrained=False, **kw):
    _no_pretrained("regnetx_160", pretrained)
    return RegNet((2, 6, 13, 1), (256, 512, 896, 2048), 128, 0.0, **kw)


def regnety_160(pretrained=False, **kw):
    _no_pretrained("regnety_160", pretrained)
    return RegNet((2, 4, 11, 1), (224, 448, 1232, 3024), 112, 0.25, **kw)


def regnety_320(pretrained=False, **kw):
    _no_pretrained("regnety_320", pretrained)
    return RegNet((2, 5, 12, 1), (232, 696, 1392, 3712), 232, 0.25, **kw)
