"""Model zoo and registry.

Parity: reference ``models/__init__.py:1-7`` (``build_model(arch, **kw)`` resolves a
constructor by name and raises ``KeyError`` for unknown names; the reference trainer then
falls back to timm, trainer.py:117-128).  Here the timm-sourced architectures the shipped
configs name (RegNetX/Y, EfficientNet-B0) are first-class members of the registry.
"""
from .botnet import *  # noqa: F401,F403
from .densenet import *  # noqa: F401,F403
from .efficientnet import *  # noqa: F401,F403
from .regnet import *  # noqa: F401,F403
from .resnet import *  # noqa: F401,F403

_REGISTRY = {name: obj for name, obj in list(globals().items())
             if callable(obj) and not name.startswith("_")}


def list_models():
    return sorted(n for n, o in _REGISTRY.items() if not isinstance(o, type))


def build_model(arch, **kwargs):
    """Instantiate ``arch``; unknown names raise ``KeyError`` like the reference."""
    return _REGISTRY[arch](**kwargs)
