"""Failure detection: step watchdog + per-rank heartbeat files.

The reference has none (no elastic agent, no watchdog, no heartbeat; a hung rank blocks the job until
NCCL's default timeout and recovery is "restart and let AUTO_RESUME pick the last epoch", trainer.py:144-146).
Here:

* ``StepWatchdog`` -- a daemon thread that expects ``tick()`` once per iteration.  If none arrives within
  ``B200.WATCHDOG_S`` seconds it logs the Python stacks of every thread (the hung collective / kernel launch shows
  up there) and, with ``B200.WATCHDOG_ABORT``, exits the process with code 3 so that torchrun / Slurm tears the job
  down and the restart resumes from the last checkpoint.  Device-side waits are bounded separately: every spin loop
  in the kernels (mbarrier waits, peer flag waits) traps after a cycle budget instead of hanging the GPU.
* ``Heartbeat`` -- each rank writes ``OUT_DIR/heartbeat/rank_<r>.json`` (epoch, iteration, wall time) every
  ``B200.HEARTBEAT_FREQ`` iterations; ``stale_ranks`` (also ``python -m distribuuuu_b200.utils.health OUT_DIR``)
  lists ranks whose last beat is older than a threshold -- enough to tell *which* node of a multi-node job died.
"""
from __future__ import annotations

import faulthandler
import json
import os
import sys
import threading
import time
from typing import Callable, List, Optional


class StepWatchdog:
    def __init__(self, timeout_s: float, abort: bool = False, on_timeout: Optional[Callable[[float], None]] = None,
                 name: str = "train"):
        self.timeout_s, self.abort, self.on_timeout, self.name = float(timeout_s), abort, on_timeout, name
        self._last = time.monotonic()
        self._stop = threading.Event()
        self._fired = 0
        self._thread: Optional[threading.Thread] = None

    @property
    def fired(self) -> int:
        return self._fired

    def tick(self) -> None:
        self._last = time.monotonic()

    def start(self) -> "StepWatchdog":
        if self.timeout_s > 0 and self._thread is None:
            self._last = time.monotonic()
            self._thread = threading.Thread(target=self._run, name=f"b200-watchdog-{self.name}", daemon=True)
            self._thread.start()
        return self

    def stop(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
            self._thread = None

    def __enter__(self):
        return self.start()

    def __exit__(self, *exc):
        self.stop()

    def _run(self) -> None:
        poll = min(max(self.timeout_s / 4.0, 0.01), 5.0)
        while not self._stop.wait(poll):
            idle = time.monotonic() - self._last
            if idle < self.timeout_s:
                continue
            self._fired += 1
            msg = (f"[b200 watchdog:{self.name}] no iteration finished for {idle:.1f}s (limit {self.timeout_s:.1f}s) "
                   f"on rank {os.environ.get('RANK', '0')}; thread stacks follow")
            try:
                from loguru import logger
                logger.error(msg)
            except Exception:  # pragma: no cover
                print(msg, file=sys.stderr, flush=True)
            try:
                faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
            except Exception:  # pragma: no cover
                pass
            if self.on_timeout is not None:
                self.on_timeout(idle)
            if self.abort:
                os._exit(3)
            self._last = time.monotonic()  # report again only after another full period


class Heartbeat:
    def __init__(self, out_dir: str, rank: int, freq: int):
        self.freq = int(freq)
        self.path = os.path.join(out_dir, "heartbeat", f"rank_{rank}.json")
        self.rank = rank
        if self.freq > 0:
            os.makedirs(os.path.dirname(self.path), exist_ok=True)

    def beat(self, epoch: int, iteration: int, force: bool = False) -> None:
        if self.freq <= 0 or not (force or iteration % self.freq == 0):
            return
        tmp = self.path + ".tmp"
        with open(tmp, "w") as f:
            json.dump({"rank": self.rank, "epoch": epoch, "iter": iteration, "time": time.time(),
                       "host": os.uname().nodename, "pid": os.getpid()}, f)
        os.replace(tmp, self.path)  # atomic: readers never see a partial file


def read_heartbeats(out_dir: str) -> List[dict]:
    d = os.path.join(out_dir, "heartbeat")
    beats = []
    if os.path.isdir(d):
        for name in sorted(os.listdir(d)):
            if name.startswith("rank_") and name.endswith(".json"):
                try:
                    beats.append(json.load(open(os.path.join(d, name))))
                except Exception:
                    pass
    return beats


def stale_ranks(out_dir: str, max_age_s: float, now: Optional[float] = None) -> List[int]:
    now = time.time() if now is None else now
    return [b["rank"] for b in read_heartbeats(out_dir) if now - b["time"] > max_age_s]


if __name__ == "__main__":  # python -m distribuuuu_b200.utils.health OUT_DIR [max_age_s]
    out = sys.argv[1] if len(sys.argv) > 1 else "."
    age = float(sys.argv[2]) if len(sys.argv) > 2 else 300.0
    for b in read_heartbeats(out):
        print(f"rank {b['rank']:4d} epoch {b['epoch']:3d} iter {b['iter']:6d} {time.time() - b['time']:8.1f}s ago on {b.get('host')}")
    dead = stale_ranks(out, age)
    print("stale ranks:", dead if dead else "none")
    sys.exit(1 if dead else 0)
