"""Epoch-granular learning-rate policies with linear warm-up.

Parity: reference ``utils.py:280-316``: ``cos`` = half-period cosine where MIN_LR is a
*fraction* of BASE_LR, ``steps`` = LR_MULT ** (index of the last boundary passed),
policy chosen by name, multiplied by BASE_LR and by a warm-up factor that rises linearly
from WARMUP_FACTOR to 1 over WARMUP_EPOCHS.
"""
from __future__ import annotations

import math

from ..config import cfg


def lr_fun_steps(cur_epoch: float) -> float:
    passed = [i for i, s in enumerate(cfg.OPTIM.STEPS) if cur_epoch >= s]
    if not passed:
        raise ValueError("OPTIM.STEPS must start at epoch 0 for the 'steps' policy")
    return cfg.OPTIM.LR_MULT ** passed[-1]


def lr_fun_cos(cur_epoch: float) -> float:
    base = 0.5 * (1.0 + math.cos(math.pi * cur_epoch / cfg.OPTIM.MAX_EPOCH))
    return (1.0 - cfg.OPTIM.MIN_LR) * base + cfg.OPTIM.MIN_LR


_POLICIES = {"cos": lr_fun_cos, "steps": lr_fun_steps}


def get_lr_fun():
    try:
        return _POLICIES[cfg.OPTIM.LR_POLICY]
    except KeyError:
        raise AssertionError("Unknown LR policy: " + str(cfg.OPTIM.LR_POLICY)) from None


def get_epoch_lr(cur_epoch: float) -> float:
    lr = get_lr_fun()(cur_epoch) * cfg.OPTIM.BASE_LR
    if cur_epoch < cfg.OPTIM.WARMUP_EPOCHS:
        alpha = cur_epoch / cfg.OPTIM.WARMUP_EPOCHS
        lr *= cfg.OPTIM.WARMUP_FACTOR * (1.0 - alpha) + alpha
    return float(lr)


def set_lr(optimizer, new_lr: float) -> None:
    for group in optimizer.param_groups:
        group["lr"] = new_lr
