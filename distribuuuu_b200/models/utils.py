"""Pretrained-weight loading shim (reference ``models/utils.py:1-4`` re-exports
``torch.hub.load_state_dict_from_url``).  This image has no network: the call still goes
through torch.hub (so a pre-populated hub cache works) and fails with an actionable
message otherwise."""
from __future__ import annotations

import torch

try:
    from torch.hub import load_state_dict_from_url
except ImportError:  # very old torch
    from torch.utils.model_zoo import load_url as load_state_dict_from_url


def load_pretrained(model: torch.nn.Module, url: str, progress: bool = True, strict: bool = True) -> None:
    try:
        state = load_state_dict_from_url(url, progress=progress, map_location="cpu")
    except Exception as exc:
        raise RuntimeError(
            f"MODEL.PRETRAINED requested but '{url}' could not be fetched ({type(exc).__name__}: {exc}). "
            "Place the file in the torch hub cache (~/.cache/torch/hub/checkpoints) or pass MODEL.WEIGHTS.") from exc
    model.load_state_dict(state, strict=strict)
