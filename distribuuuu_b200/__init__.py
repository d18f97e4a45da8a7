"""distribuuuu_b200 -- Blackwell-native distributed image-classification training.

Same capabilities as BIGBALLON/distribuuuu (entry points, config schema, model zoo,
SyncBN, SGD/cosine recipe, checkpoint layout, launch modes); the compute path is
hand-written sm_100a CUDA (``csrc/``) driven from PyTorch, with a peer-memory
gradient all-reduce fused with the optimizer update.
"""
__version__ = "0.1.0"

from . import config, models, ops, parallel, trainer, utils  # noqa: F401
from .config import cfg  # noqa: F401
