"""ResNet / ResNeXt / Wide-ResNet (v1.5: the stride sits on the 3x3 of a bottleneck).

Parity: reference ``distribuuuu/models/resnet.py`` (blocks 57-161, trunk 164-297,
constructors 315-447).  Parameter/buffer names follow torchvision (``conv1``, ``bn1``,
``layerK.i.convJ``, ``layerK.i.downsample.0/1``, ``fc``) so checkpoints interchange.
The graph is expressed through ``ops.functional`` so that the same definition runs as
plain torch ops (CPU / reference semantics) or as fused sm_100a kernels (native engine).
"""
from __future__ import annotations

import torch.nn as nn

from ..ops import functional as Fn
from .utils import load_pretrained

__all__ = ["ResNet", "BasicBlock", "Bottleneck", "resnet18", "resnet34", "resnet50", "resnet101",
           "resnet152", "resnext50_32x4d", "resnext101_32x8d", "wide_resnet50_2", "wide_resnet101_2"]

_HUB = "https://download.pytorch.org/models/"

# arch -> (block name, blocks per stage, constructor overrides, torchvision checkpoint file)
# (the variants of reference models/resnet.py:315-447 and its URL table :23-33, as data)
_SPECS = {
    "resnet18": ("BasicBlock", (2, 2, 2, 2), {}, "resnet18-5c106cde.pth"),
    "resnet34": ("BasicBlock", (3, 4, 6, 3), {}, "resnet34-333f7ec4.pth"),
    "resnet50": ("Bottleneck", (3, 4, 6, 3), {}, "resnet50-19c8e357.pth"),
    "resnet101": ("Bottleneck", (3, 4, 23, 3), {}, "resnet101-5d3b4d8f.pth"),
    "resnet152": ("Bottleneck", (3, 8, 36, 3), {}, "resnet152-b121ed2d.pth"),
    "resnext50_32x4d": ("Bottleneck", (3, 4, 6, 3), {"groups": 32, "width_per_group": 4}, "resnext50_32x4d-7cdf4587.pth"),
    "resnext101_32x8d": ("Bottleneck", (3, 4, 23, 3), {"groups": 32, "width_per_group": 8}, "resnext101_32x8d-8ba56ff5.pth"),
    "wide_resnet50_2": ("Bottleneck", (3, 4, 6, 3), {"width_per_group": 128}, "wide_resnet50_2-95faca4d.pth"),
    "wide_resnet101_2": ("Bottleneck", (3, 4, 23, 3), {"width_per_group": 128}, "wide_resnet101_2-32ee1156.pth"),
}
model_urls = {arch: _HUB + spec[3] for arch, spec in _SPECS.items()}


def conv3x3(cin: int, cout: int, stride: int = 1, groups: int = 1, dilation: int = 1) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, 3, stride=stride, padding=dilation, groups=groups, bias=False, dilation=dilation)


def conv1x1(cin: int, cout: int, stride: int = 1) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, 1, stride=stride, bias=False)


class _Shortcut(nn.Sequential):
    """1x1 projection + BN; indices 0/1 give the torchvision ``downsample.0/1`` keys."""

    def __init__(self, cin, cout, stride, norm_layer):
        super().__init__(conv1x1(cin, cout, stride), norm_layer(cout))

    def forward(self, x):
        return Fn.conv_bn_act(x, self[0], self[1], None)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64,
                 dilation=1, norm_layer=None):
        super().__init__()
        norm_layer = norm_layer or nn.BatchNorm2d
        if groups != 1 or base_width != 64:
            raise ValueError("BasicBlock only supports groups=1 and base_width=64")
        if dilation > 1:
            raise NotImplementedError("Dilation > 1 not supported in BasicBlock")
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = norm_layer(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = norm_layer(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        # projection shortcut: its conv read x first, so conv1's input gradient is folded into that conv's dgrad
        proj = self.downsample[0] if isinstance(self.downsample, _Shortcut) else None
        out = Fn.conv_bn_act(x, self.conv1, self.bn1, "relu", input_grad_to=proj)
        sink = self.conv1 if self.downsample is None else None   # identity shortcut: conv1 reads the same tensor
        return Fn.conv_bn_act(out, self.conv2, self.bn2, "relu", residual=identity, residual_sink=sink)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64,
                 dilation=1, norm_layer=None):
        super().__init__()
        norm_layer = norm_layer or nn.BatchNorm2d
        width = int(planes * (base_width / 64.0)) * groups
        self.conv1 = conv1x1(inplanes, width)
        self.bn1 = norm_layer(width)
        self.conv2 = conv3x3(width, width, stride, groups, dilation)
        self.bn2 = norm_layer(width)
        self.conv3 = conv1x1(width, planes * self.expansion)
        self.bn3 = norm_layer(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        # projection shortcut: its conv read x first, so conv1's input gradient is folded into that conv's dgrad
        proj = self.downsample[0] if isinstance(self.downsample, _Shortcut) else None
        out = Fn.conv_bn_act(x, self.conv1, self.bn1, "relu", input_grad_to=proj)
        out = Fn.conv_bn_act(out, self.conv2, self.bn2, "relu")
        sink = self.conv1 if self.downsample is None else None   # identity shortcut: conv1 reads the same tensor
        return Fn.conv_bn_act(out, self.conv3, self.bn3, "relu", residual=identity, residual_sink=sink)


class ResNet(nn.Module):
    def __init__(self, block, layers, num_classes=1000, zero_init_residual=False, groups=1,
                 width_per_group=64, replace_stride_with_dilation=None, norm_layer=None):
        super().__init__()
        self._norm_layer = norm_layer or nn.BatchNorm2d
        self.inplanes, self.dilation = 64, 1
        self.groups, self.base_width = groups, width_per_group
        rswd = replace_stride_with_dilation or [False, False, False]
        if len(rswd) != 3:
            raise ValueError(f"replace_stride_with_dilation should be None or a 3-element tuple, got {rswd}")

        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = self._norm_layer(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2, dilate=rswd[0])
        self.layer3 = self._make_layer(block, 256, layers[2], stride=2, dilate=rswd[1])
        self.layer4 = self._make_layer(block, 512, layers[3], stride=2, dilate=rswd[2])
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, num_classes)

        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.zeros_(m.bn3.weight)
                elif isinstance(m, BasicBlock):
                    nn.init.zeros_(m.bn2.weight)

    def _make_layer(self, block, planes, blocks, stride=1, dilate=False):
        prev_dilation = self.dilation
        if dilate:
            self.dilation *= stride
            stride = 1
        shortcut = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            shortcut = _Shortcut(self.inplanes, planes * block.expansion, stride, self._norm_layer)
        stage = [block(self.inplanes, planes, stride, shortcut, self.groups, self.base_width,
                       prev_dilation, self._norm_layer)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            stage.append(block(self.inplanes, planes, groups=self.groups, base_width=self.base_width,
                               dilation=self.dilation, norm_layer=self._norm_layer))
        return nn.Sequential(*stage)

    def forward_stem(self, x):
        return Fn.conv_bn_relu_maxpool(x, self.conv1, self.bn1, 3, 2, 1)

    def forward_features(self, x):
        x = self.forward_stem(x)
        x = self.layer1(x)
        x = self.layer2(x)
        x = self.layer3(x)
        return self.layer4(x)

    def forward(self, x):
        x = self.forward_features(x)
        return Fn.linear(Fn.global_avg_pool(x), self.fc)


def _make_constructor(arch: str):
    block_name, layers, overrides, _ = _SPECS[arch]

    def build(pretrained: bool = False, progress: bool = True, **kwargs) -> ResNet:
        kwargs.update(overrides)
        model = ResNet(globals()[block_name], list(layers), **kwargs)
        if pretrained:
            load_pretrained(model, model_urls[arch], progress)
        return model

    build.__name__ = build.__qualname__ = arch
    build.__doc__ = (f"{arch}: {block_name} x {layers}" + (f", {overrides}" if overrides else "") +
                     "; ``pretrained`` loads torchvision's ImageNet weights (needs network access).")
    return build


for _arch in _SPECS:
    globals()[_arch] = _make_constructor(_arch)
del _arch
