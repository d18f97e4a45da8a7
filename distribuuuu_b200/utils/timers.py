"""CUDA-event phase timers (absent from the reference, SURVEY 5 'tracing/profiling').

``StepTimer`` brackets arbitrary regions with CUDA events on the current stream (or
``time.perf_counter`` on CPU) and reports per-phase milliseconds without forcing a host
sync until ``summary`` is called.
"""
from __future__ import annotations

import time
from collections import defaultdict

import torch


class StepTimer:
    def __init__(self, device: torch.device, enabled: bool = True):
        self.cuda = device.type == "cuda"
        self.enabled = enabled
        self._open = {}
        self._pairs = defaultdict(list)

    def start(self, name: str) -> None:
        if not self.enabled:
            return
        if self.cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._open[name] = ev
        else:
            self._open[name] = time.perf_counter()

    def stop(self, name: str) -> None:
        if not self.enabled or name not in self._open:
            return
        begin = self._open.pop(name)
        if self.cuda:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            self._pairs[name].append((begin, end))
        else:
            self._pairs[name].append((begin, time.perf_counter()))

    def summary(self, reset: bool = True) -> dict:
        if self.cuda:
            torch.cuda.synchronize()
        out = {}
        for name, pairs in self._pairs.items():
            if self.cuda:
                ms = [b.elapsed_time(e) for b, e in pairs]
            else:
                ms = [(e - b) * 1e3 for b, e in pairs]
            out[name] = {"n": len(ms), "mean_ms": sum(ms) / max(len(ms), 1), "total_ms": sum(ms)}
        if reset:
            self._pairs.clear()
        return out
