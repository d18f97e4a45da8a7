"""DenseNet-121/161/169/201.

Parity: reference ``distribuuuu/models/densenet.py`` (dense layer 23-117 incl. the
``memory_efficient`` activation-checkpointing flag, block 120-148, transition 151-166,
trunk 169-263, legacy-key remap 266-282, constructors 300-365).  Key names match
torchvision (``features.denseblockK.denselayerJ.norm1`` ...).

B200 note: a dense block owns one pre-allocated NHWC buffer (``Fn.dense_block_buffer``); every layer's 3x3 conv
TMA-stores its new features into the next channel slice and ``Fn.concat_channels`` of adjacent slices is a view, so
the per-layer ``torch.cat`` of the reference (densenet.py:68,148) moves no bytes on the native path.
"""
from __future__ import annotations

import functools
import re
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.utils.checkpoint as cp

from ..ops import functional as Fn
from ..ops import runtime
from .utils import load_state_dict_from_url

__all__ = ["DenseNet", "densenet121", "densenet169", "densenet201", "densenet161"]

model_urls = {
    "densenet121": "https://download.pytorch.org/models/densenet121-a639ec97.pth",
    "densenet169": "https://download.pytorch.org/models/densenet169-b2777c0a.pth",
    "densenet201": "https://download.pytorch.org/models/densenet201-c1103571.pth",
    "densenet161": "https://download.pytorch.org/models/densenet161-8d451a50.pth",
}


class _DenseLayer(nn.Module):
    """BN-ReLU-1x1(bn_size*k) -> BN-ReLU-3x3(k), optional dropout / recomputation."""

    def __init__(self, cin: int, growth_rate: int, bn_size: int, drop_rate: float, memory_efficient: bool = False):
        super().__init__()
        self.norm1 = nn.BatchNorm2d(cin)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv1 = nn.Conv2d(cin, bn_size * growth_rate, 1, bias=False)
        self.norm2 = nn.BatchNorm2d(bn_size * growth_rate)
        self.relu2 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(bn_size * growth_rate, growth_rate, 3, padding=1, bias=False)
        self.drop_rate = float(drop_rate)
        self.memory_efficient = memory_efficient

    def _bottleneck(self, *features):
        x = Fn.concat_channels(features)
        return Fn.conv2d(Fn.bn_act(x, self.norm1, "relu"), self.conv1)

    def _bottleneck_on(self, engine, *features):
        # The recomputation runs during backward -- outside the engine's forward scope and, on CUDA, on autograd's
        # worker thread -- so the execution path chosen in the forward pass is re-entered explicitly; otherwise the
        # recomputed graph (ATen ops) would not match the recorded one (native ops).
        with runtime.native_scope(engine):
            return self._bottleneck(*features)

    def forward(self, features, out_buffer=None):
        feats = [features] if torch.is_tensor(features) else list(features)
        if self.memory_efficient and any(f.requires_grad for f in feats):
            mid = cp.checkpoint(functools.partial(self._bottleneck_on, runtime.active_engine()), *feats, use_reentrant=False)
        else:
            mid = self._bottleneck(*feats)
        new = Fn.conv2d(Fn.bn_act(mid, self.norm2, "relu"), self.conv2, out_buffer=out_buffer)
        return Fn.dropout(new, self.drop_rate, self.training)


class _DenseBlock(nn.ModuleDict):
    def __init__(self, num_layers, cin, bn_size, growth_rate, drop_rate, memory_efficient=False):
        super().__init__()
        for i in range(num_layers):
            self[f"denselayer{i + 1}"] = _DenseLayer(cin + i * growth_rate, growth_rate, bn_size,
                                                     drop_rate, memory_efficient)
        self.total_channels = cin + num_layers * growth_rate

    def forward(self, x):
        x, buf = Fn.dense_block_buffer(x, self.total_channels)
        feats = [x]
        for layer in self.values():
            feats.append(layer(feats, buf))
        return Fn.concat_channels(feats)


class _Transition(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.norm = nn.BatchNorm2d(cin)
        self.relu = nn.ReLU(inplace=True)
        self.conv = nn.Conv2d(cin, cout, 1, bias=False)
        self.pool = nn.AvgPool2d(2, 2)

    def forward(self, x):
        x = Fn.conv2d(Fn.bn_act(x, self.norm, "relu"), self.conv)
        return Fn.avg_pool2d(x, 2, 2)


class _Features(nn.Module):
    """Holds conv0/norm0/.../norm5 under the torchvision names and runs them in order."""

    def forward(self, x):
        x = Fn.conv_bn_relu_maxpool(x, self.conv0, self.norm0, 3, 2, 1)
        for name, mod in self.named_children():
            if name.startswith(("denseblock", "transition")):
                x = mod(x)
        return Fn.bn_act(x, self.norm5, "relu")


class DenseNet(nn.Module):
    def __init__(self, growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4,
                 drop_rate=0.0, num_classes=1000, memory_efficient=False):
        super().__init__()
        f = _Features()
        f.add_module("conv0", nn.Conv2d(3, num_init_features, 7, stride=2, padding=3, bias=False))
        f.add_module("norm0", nn.BatchNorm2d(num_init_features))
        f.add_module("relu0", nn.ReLU(inplace=True))
        f.add_module("pool0", nn.MaxPool2d(3, stride=2, padding=1))
        width = num_init_features
        for i, n in enumerate(block_config):
            f.add_module(f"denseblock{i + 1}", _DenseBlock(n, width, bn_size, growth_rate, drop_rate, memory_efficient))
            width += n * growth_rate
            if i != len(block_config) - 1:
                f.add_module(f"transition{i + 1}", _Transition(width, width // 2))
                width //= 2
        f.add_module("norm5", nn.BatchNorm2d(width))
        self.features = f
        self.classifier = nn.Linear(width, num_classes)

        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.zeros_(m.bias)

    def forward(self, x):
        x = self.features(x)
        return Fn.linear(Fn.global_avg_pool(x), self.classifier)


_LEGACY_PART = re.compile(r"(denselayer\d+\.(?:norm|relu|conv))\.([12])\.")


def _load_state_dict(model, url, progress):
    """Old torchvision checkpoints say 'norm.1.weight' / 'conv.2.weight'; the dot was dropped
    later ('norm1.weight').  Collapse it so both generations load."""
    state = load_state_dict_from_url(url, progress=progress, map_location="cpu")
    model.load_state_dict(OrderedDict((_LEGACY_PART.sub(r"\1\2.", k), v) for k, v in state.items()))


def _densenet(arch, growth_rate, block_config, num_init_features, pretrained, progress, **kw):
    model = DenseNet(growth_rate, block_config, num_init_features, **kw)
    if pretrained:
        _load_state_dict(model, model_urls[arch], progress)
    return model


def densenet121(pretrained=False, progress=True, **kw):
    return _densenet("densenet121", 32, (6, 12, 24, 16), 64, pretrained, progress, **kw)


def densenet161(pretrained=False, progress=True, **kw):
    return _densenet("densenet161", 48, (6, 12, 36, 24), 96, pretrained, progress, **kw)


def densenet169(pretrained=False, progress=True, **kw):
    return _densenet("densenet169", 32, (6, 12, 32, 32), 64, pretrained, progress, **kw)


def densenet201(pretrained=False, progress=True, **kw):
    return _densenet("densenet201", 32, (6, 12, 48, 32), 64, pretrained, progress, **kw)
