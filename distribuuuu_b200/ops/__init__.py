"""Operator layer: functional API (``functional``), path selection (``runtime``) and the
sm_100a kernel bindings (``native``, imported lazily so CPU-only use never needs nvcc)."""
from . import functional, runtime  # noqa: F401
