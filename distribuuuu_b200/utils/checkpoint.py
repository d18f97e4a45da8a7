"""Checkpoint layout, save and (auto-)resume.

Parity: reference ``utils.py:319-350`` (``OUT_DIR/checkpoints/ckpt_ep_NNN.pth.tar``,
"last" = lexicographic max) and ``utils.py:360-410`` (rank 0 writes
``{"epoch", "state_dict", "optimizer", "best_acc1"}`` with un-prefixed module keys and a
weights-only ``OUT_DIR/best.pth.tar``; load accepts a full dict or a bare state_dict,
restores epoch/best only if the optimizer state loaded).

The optimizer entry is always in ``torch.optim.SGD.state_dict()`` format, including when
the fused flat optimizer of the native engine produced it, so files interchange with the
reference in both directions.
"""
from __future__ import annotations

import os

import torch
from loguru import logger

from ..config import cfg
from .dist import get_rank

_NAME_PREFIX = "ckpt_ep_"
_DIR_NAME = "checkpoints"


def get_checkpoint_dir() -> str:
    return os.path.join(cfg.OUT_DIR, _DIR_NAME)


def get_checkpoint(epoch: int) -> str:
    return os.path.join(get_checkpoint_dir(), f"{_NAME_PREFIX}{epoch:03d}.pth.tar")


def _list_checkpoints():
    d = get_checkpoint_dir()
    if not os.path.isdir(d):
        return []
    # complete files only: a crash during a save leaves ``.<name>.tmp`` behind, which must never be resumed from
    return sorted(f for f in os.listdir(d) if f.startswith(_NAME_PREFIX) and f.endswith(".pth.tar"))


def get_last_checkpoint() -> str:
    names = _list_checkpoints()
    if not names:
        raise FileNotFoundError(f"no checkpoints under {get_checkpoint_dir()}")
    return os.path.join(get_checkpoint_dir(), names[-1])


def has_checkpoint() -> bool:
    return bool(_list_checkpoints())


def unwrap_model(model):
    """Strip any data-parallel wrapper (torch DDP/DP or this package's engines)."""
    while hasattr(model, "module") and isinstance(getattr(model, "module"), torch.nn.Module):
        model = model.module
    return model


def count_parameters(model) -> str:
    n = sum(p.numel() for p in model.parameters() if p.requires_grad)
    return f"Params(M): {n / 1e6:.3f}, Model Size(MB): {n * 4.0 / 1024 / 1024:.3f}"


def _cpu_state_dict(model) -> dict:
    return {k: v.detach().to("cpu", copy=True).contiguous() if torch.is_tensor(v) else v
            for k, v in unwrap_model(model).state_dict().items()}


def _atomic_save(obj, path: str):
    """Write to a hidden temporary name (no ``ckpt_ep_`` prefix, so AUTO_RESUME never sees it) and rename."""
    d, name = os.path.split(path)
    tmp = os.path.join(d, f".{name}.tmp")
    torch.save(obj, tmp)
    os.replace(tmp, path)


def save_checkpoint(model, optimizer, epoch: int, best_acc1: float, best: bool):
    """Rank 0 writes ``ckpt_ep_{epoch+1:03d}`` (+ ``best.pth.tar``); others return None.

    ``optimizer.state_dict()`` may be collective for sharded optimizers, so it is
    called on every rank before the rank check.
    """
    opt_state = optimizer.state_dict() if optimizer is not None else None
    if get_rank() != 0:
        return None
    os.makedirs(get_checkpoint_dir(), exist_ok=True)
    state = _cpu_state_dict(model)
    path = get_checkpoint(epoch + 1)
    _atomic_save({"epoch": epoch, "state_dict": state, "optimizer": opt_state, "best_acc1": best_acc1}, path)
    if best:
        _atomic_save(state, os.path.join(cfg.OUT_DIR, "best.pth.tar"))
    return path


def load_checkpoint(checkpoint_file: str, model, optimizer=None):
    """Load weights (and optionally optimizer state); returns (start_epoch, best_acc1)."""
    assert os.path.exists(checkpoint_file), f"CHECKPOINT '{checkpoint_file}' NOT FOUND"
    start_epoch, best_acc1 = 0, 0
    ckpt = torch.load(checkpoint_file, map_location="cpu", weights_only=False)
    target = unwrap_model(model)
    if isinstance(ckpt, dict) and "state_dict" in ckpt:
        target.load_state_dict(ckpt["state_dict"])
        if optimizer is not None:
            try:
                optimizer.load_state_dict(ckpt["optimizer"])
                start_epoch = ckpt["epoch"] + 1
                best_acc1 = ckpt["best_acc1"]
            except Exception as exc:  # same tolerance as the reference, but say why
                logger.info(f"CAN'T FOUND OPTIMIZER in {checkpoint_file} ({type(exc).__name__}: {exc})")
    else:
        target.load_state_dict(ckpt)
    if hasattr(model, "on_weights_loaded"):
        model.on_weights_loaded()
    if get_rank() == 0:
        logger.info(f"LOADED '{checkpoint_file}'")
    return start_epoch, best_acc1
