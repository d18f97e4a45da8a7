"""Optimizer factory.

Parity: reference ``utils.py:187-196`` -- Nesterov SGD over a single parameter group
(weight decay on every parameter, BN and bias included).  On the reference-semantics
path this is ``torch.optim.SGD``; the native engine substitutes
``parallel.native_engine.FusedSGD``, which exposes the same ``param_groups`` /
``state_dict`` / ``load_state_dict`` surface.
"""
from __future__ import annotations

import torch

from ..config import cfg


def sgd_hparams() -> dict:
    return dict(lr=cfg.OPTIM.BASE_LR, momentum=cfg.OPTIM.MOMENTUM, weight_decay=cfg.OPTIM.WEIGHT_DECAY,
                dampening=cfg.OPTIM.DAMPENING, nesterov=cfg.OPTIM.NESTEROV)


def construct_optimizer(model):
    if hasattr(model, "make_optimizer"):  # engines that own a fused optimizer
        return model.make_optimizer(**sgd_hparams())
    return torch.optim.SGD(model.parameters(), **sgd_hparams())
