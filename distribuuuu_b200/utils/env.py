"""Seeding, output directory and logging setup.

Parity: reference ``utils.py:54-68`` (``setup_seed``: rank 0 creates OUT_DIR and dumps
the config; a truthy RNG_SEED seeds numpy/torch/random with ``seed + rank`` and forces
deterministic cuDNN, otherwise the CUDNN.* flags apply) and ``utils.py:71-82``
(``setup_logger``: loguru, file sink on rank 0, stderr on every rank,
``[YYYY-MM-DD HH:mm:ss] message``).
"""
from __future__ import annotations

import os
import random
import sys
import time

import numpy as np
import torch
from loguru import logger

from .. import config
from ..config import cfg

_FMT = "[{time:YYYY-MM-DD HH:mm:ss}] {message}"


def setup_seed(rank: int) -> None:
    if rank == 0:
        os.makedirs(cfg.OUT_DIR, exist_ok=True)
        config.dump_cfg()
    if cfg.RNG_SEED:
        seed = int(cfg.RNG_SEED) + rank
        np.random.seed(seed)
        torch.manual_seed(seed)
        random.seed(seed)
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.deterministic = True
    else:
        torch.backends.cudnn.benchmark = bool(cfg.CUDNN.BENCHMARK)
        torch.backends.cudnn.deterministic = bool(cfg.CUDNN.DETERMINISTIC)


def setup_logger(rank: int, local_rank: int) -> None:
    logger.remove()
    if rank == 0:
        os.makedirs(cfg.OUT_DIR, exist_ok=True)
        logger.add(os.path.join(cfg.OUT_DIR, f"{time.time()}.log"), format=_FMT)
    logger.add(sys.stderr, format=_FMT)
    logger.debug(f"LOCAL_RANK: {local_rank}, RANK: {rank}")
    if rank == 0:
        logger.debug(f"\n{cfg.dump()}")
