"""Command-line front end shared by ``train_net.py`` and ``test_net.py``.

Contract (reference train_net.py:6-9 / test_net.py:6-9): ``--cfg FILE`` then ``KEY VALUE`` overrides, the config
is frozen before any work starts, ``train`` runs the epoch loop and ``test`` a single validation pass.  On top of
that the process group is shut down on every exit path, so an exception on one rank does not leave its peers
blocked in a collective until the timeout.
"""
from __future__ import annotations

import sys
from typing import Optional, Sequence

_MODES = {"train": ("Train a classification model.", "train_model"),
          "test": ("Test a classification model.", "test_model")}


def run(mode: str, argv: Optional[Sequence[str]] = None):
    """Parse the command line, freeze the config and run ``mode`` (``"train"`` or ``"test"``); returns whatever the
    trainer entry returns (best top-1 for training, ``(top1, topk)`` for evaluation)."""
    from . import config, trainer, utils
    description, entry = _MODES[mode]
    config.load_cfg_fom_args(description, argv=argv)
    config.cfg.freeze()
    try:
        result = getattr(trainer, entry)()
        utils.barrier()          # success path: nobody tears the group down while a peer is still reducing metrics
        return result
    finally:
        utils.shutdown()


def main(mode: str) -> None:
    run(mode, sys.argv[1:])
