"""Execution-path selection for the functional ops.

Two paths share one model definition:

* ``torch``  -- reference semantics on any device (plain ATen ops, fp32 by default);
  this is what the CPU/gloo plumbing configuration runs.
* ``native`` -- sm_100a kernels on NHWC bf16 activations.  Active only inside
  ``native_scope(engine)``, which the native data-parallel engine enters around the
  model's forward; the engine owns the bf16 compute copies of the parameters and the
  flat gradient buckets.
"""
from __future__ import annotations

import contextlib
import threading

_state = threading.local()


def active_engine():
    """The native engine whose forward is currently executing, or None."""
    return getattr(_state, "engine", None)


@contextlib.contextmanager
def native_scope(engine):
    prev = getattr(_state, "engine", None)
    _state.engine = engine
    try:
        yield engine
    finally:
        _state.engine = prev
