"""Running meters, progress line and top-k accuracy.

Parity: reference ``utils.py:199-262`` (AverageMeter / ProgressMeter / construct_meters,
log line ``ETA: x.xxh PREFIX[ i/N] | Time v (avg) | Data .. | Loss .. | Acc@1 .. | Acc@k ..``)
and ``utils.py:265-277`` (accuracy in percent, one 1-element tensor per k).
``DeviceMetrics`` is new: it keeps loss / hit counts on the device so the host only
synchronises every ``B200.METRIC_SYNC_FREQ`` iterations.
"""
from __future__ import annotations

import time

import torch
from loguru import logger

from ..config import cfg


class AverageMeter:
    """Tracks the latest value, the weighted sum, the count and the running mean."""

    def __init__(self, name: str, fmt: str = ":f"):
        self.name, self.fmt = name, fmt
        self.reset()

    def reset(self) -> None:
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n: int = 1) -> None:
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / max(self.count, 1)

    def __str__(self) -> str:
        return ("{name} {val" + self.fmt + "} ({avg" + self.fmt + "})").format(
            name=self.name, val=self.val, avg=self.avg)


class ProgressMeter:
    """Formats one progress line per call and estimates time to completion."""

    def __init__(self, num_batches: int, meters, prefix: str = ""):
        width = len(str(int(num_batches)))
        self._batch_fmt = "[{:" + str(width) + "d}/" + str(int(num_batches)) + "]"
        self.meters = list(meters)
        self.prefix = prefix
        self.time_eta = None

    def display(self, batch: int) -> None:
        line = " | ".join([self.prefix + self._batch_fmt.format(batch)] + [str(m) for m in self.meters])
        if self.time_eta:
            line = f"ETA: {self.time_eta / 3600:.2f}h " + line
        logger.info(line)

    def cal_eta(self, iters: int, total_iter: int, tic: float | None = None,
                cur_epoch: int = 0, start_epoch: int = 0) -> None:
        elapsed = time.time() - (tic if tic is not None else time.time())
        max_epoch = cfg.OPTIM.MAX_EPOCH
        done_this_run = (cur_epoch - start_epoch + iters / total_iter) / max_epoch
        remaining = 1 - (cur_epoch + iters / total_iter) / max_epoch
        self.time_eta = elapsed / max(done_this_run, 1e-12) * remaining


def construct_meters():
    return (AverageMeter("Time", ":6.3f"), AverageMeter("Data", ":5.3f"),
            AverageMeter("Loss", ":6.4f"), AverageMeter("Acc@1", ":6.3f"),
            AverageMeter(f"Acc@{cfg.TRAIN.TOPK}", ":6.3f"))


def accuracy(output: torch.Tensor, target: torch.Tensor, topk=(1,)):
    """Top-k accuracy in percent for each k (list of 1-element tensors)."""
    with torch.no_grad():
        kmax = min(max(topk), output.size(1))
        top = output.float().topk(kmax, dim=1).indices  # [B, kmax]
        hit = top.eq(target.view(-1, 1))
        scale = 100.0 / target.size(0)
        return [hit[:, :min(k, kmax)].any(dim=1).float().sum().reshape(1) * scale for k in topk]


class DeviceMetrics:
    """Accumulates (loss*n, top1 hits, topk hits, n) on the device.

    ``update`` is sync-free; ``flush`` does one packed all-reduce and one D2H copy and
    returns the job-wide means since the previous flush.
    """

    def __init__(self, device: torch.device):
        self.acc = torch.zeros(4, dtype=torch.float64 if device.type == "cpu" else torch.float32, device=device)

    def update(self, loss: torch.Tensor, hits1: torch.Tensor, hitsk: torch.Tensor, n: int) -> None:
        self.acc[0] += loss.detach().float() * n
        self.acc[1] += hits1.detach().float().reshape(())
        self.acc[2] += hitsk.detach().float().reshape(())
        self.acc[3] += n

    def flush(self):
        from .dist import get_world_size
        import torch.distributed as dist
        buf = self.acc.clone()
        if get_world_size() > 1:
            dist.all_reduce(buf)
        loss_sum, h1, hk, n = buf.tolist()
        self.acc.zero_()
        n = max(n, 1.0)
        return loss_sum / n, 100.0 * h1 / n, 100.0 * hk / n, int(n)
