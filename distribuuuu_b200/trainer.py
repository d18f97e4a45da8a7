"""Training / evaluation orchestration.

Parity: reference ``distribuuuu/trainer.py`` -- ``train_epoch`` 14-64, ``validate``
67-103, ``train_model`` 106-173, ``test_model`` 176-209 (same call order: distributed
init -> seed/out-dir -> logger -> model -> optional SyncBN -> device -> data-parallel
wrap -> loaders -> loss/optimizer -> resume -> epoch loop {train, validate, checkpoint,
log}).  Log line formats are kept.

What is different, on purpose:
* the data-parallel wrapper is an *engine* object with a ``train_step``: the native
  engine (CUDA) runs fused sm_100a kernels and a peer-memory all-reduce fused with the
  SGD update; the torch engine reproduces DDP semantics on any backend;
* loss / accuracy are accumulated on the device and read back once per
  ``B200.METRIC_SYNC_FREQ`` iterations in a single packed reduction (the reference does
  3 all-reduces + 3 ``.item()`` syncs every iteration);
* inputs are prefetched to the device on a side stream; timings can be CUDA-event based;
* optional failure detection (``B200.WATCHDOG_S``, ``B200.HEARTBEAT_FREQ``; utils/health.py).
"""
from __future__ import annotations

import os
import time

import torch
import torch.nn as nn
from loguru import logger

from . import models, utils
from .config import cfg
from .ops import functional as Fn
from .parallel import BucketedDataParallel, SyncBatchNorm


# ---------------------------------------------------------------------------------
# engines
# ---------------------------------------------------------------------------------
class TorchEngine(BucketedDataParallel):
    """Reference-semantics step: forward, CE, zero_grad, backward (+bucketed all-reduce),
    optimizer.step -- the sequence at reference trainer.py:42-47."""

    @staticmethod
    def _prep(inputs):
        # B200.INPUT_UINT8 batches: ToTensor + Normalize happen here instead of in the loader workers
        return utils.normalize_uint8(inputs) if inputs.dtype == torch.uint8 else inputs

    def train_step(self, inputs, targets, optimizer, topk: int):
        outputs = self(self._prep(inputs))
        loss, hits1, hitsk = Fn.cross_entropy_topk(outputs, targets, topk)
        optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self.finish_backward()
        optimizer.step()
        return loss.detach(), hits1, hitsk

    @torch.no_grad()
    def eval_step(self, inputs, targets, topk: int):
        outputs = self(self._prep(inputs))
        return Fn.cross_entropy_topk(outputs, targets, topk)


def _select_engine(device: torch.device) -> str:
    choice = cfg.B200.ENGINE
    precision = str(cfg.B200.PRECISION).lower()
    if precision not in ("bf16", "fp32"):
        raise ValueError(f"B200.PRECISION must be 'bf16' or 'fp32', got {cfg.B200.PRECISION!r}")
    if choice == "auto":
        # the native kernels compute in bf16 (fp32 masters); fp32 compute is the reference-semantics torch path
        choice = "native" if (device.type == "cuda" and precision == "bf16") else "torch"
    if choice == "native" and device.type != "cuda":
        raise RuntimeError("B200.ENGINE=native requires a CUDA device")
    if choice == "native" and precision != "bf16":
        raise ValueError("B200.ENGINE=native computes in bf16; use B200.PRECISION=bf16, or B200.ENGINE=torch/auto "
                         "for fp32 compute")
    return choice


def build_engine(net: nn.Module, device: torch.device):
    """Wrap ``net`` in the engine selected by ``cfg.B200.ENGINE``."""
    if _select_engine(device) == "native":
        from .parallel.native_engine import NativeEngine
        return NativeEngine(net, device, precision=cfg.B200.PRECISION, comm=cfg.B200.COMM,
                            bucket_cap_mb=cfg.B200.BUCKET_MB, sync_bn=cfg.MODEL.SYNCBN, cuda_graph=cfg.B200.CUDA_GRAPH)
    return TorchEngine(net, bucket_cap_mb=cfg.B200.BUCKET_MB)


def build_net():
    """Zoo lookup (reference trainer.py:117-128; the timm fallback is folded into the zoo)."""
    try:
        return models.build_model(arch=cfg.MODEL.ARCH, pretrained=cfg.MODEL.PRETRAINED,
                                  num_classes=cfg.MODEL.NUM_CLASSES)
    except KeyError:
        raise KeyError(f"unknown MODEL.ARCH '{cfg.MODEL.ARCH}'; available: {models.list_models()}") from None


def _wrap_loader(loader, device):
    if isinstance(loader, utils.SyntheticDeviceLoader):
        return loader
    return utils.PinnedPrefetcher(loader, device)


# ---------------------------------------------------------------------------------
# epoch loops
# ---------------------------------------------------------------------------------
def train_epoch(train_loader, engine, optimizer, cur_epoch, start_epoch, tic, device=None):
    """One epoch; LR is set once per epoch (reference trainer.py:25-26)."""
    rank = utils.get_rank()
    device = device or next(engine.parameters()).device
    batch_time, data_time, losses, top1, topk = utils.construct_meters()
    n_iters = len(train_loader) if not cfg.B200.MAX_ITERS else min(len(train_loader), cfg.B200.MAX_ITERS)
    progress = utils.ProgressMeter(n_iters, [batch_time, data_time, losses, top1, topk],
                                   prefix=f"TRAIN:  [{cur_epoch + 1}]")
    lr = utils.get_epoch_lr(cur_epoch)
    utils.set_lr(optimizer, lr)
    if rank == 0:
        logger.debug(f"CURRENT EPOCH: {cur_epoch + 1:3d},   LR: {lr:.4f},   POLICY: {cfg.OPTIM.LR_POLICY}")
    train_loader.sampler.set_epoch(cur_epoch)
    engine.train()

    metrics = utils.DeviceMetrics(device)
    sync_freq = max(int(cfg.B200.METRIC_SYNC_FREQ), 1)
    timer = utils.StepTimer(device, enabled=bool(cfg.B200.PROFILE))   # CUDA-event phase timing (B200.PROFILE)
    watchdog = utils.StepWatchdog(cfg.B200.WATCHDOG_S, abort=cfg.B200.WATCHDOG_ABORT, name=f"epoch{cur_epoch + 1}").start()
    heartbeat = utils.Heartbeat(cfg.OUT_DIR, rank, cfg.B200.HEARTBEAT_FREQ)
    end = time.time()
    for idx, (inputs, targets) in enumerate(train_loader):
        if idx >= n_iters:
            break
        data_time.update(time.time() - end)
        timer.start("step")
        loss, hits1, hitsk = engine.train_step(inputs, targets, optimizer, cfg.TRAIN.TOPK)
        timer.stop("step")
        metrics.update(loss, hits1, hitsk, targets.size(0))
        watchdog.tick()
        heartbeat.beat(cur_epoch, idx + 1)

        last = (idx + 1) == n_iters
        if (idx + 1) % sync_freq == 0 or last:
            m_loss, m_top1, m_topk, n = metrics.flush()  # one packed all-reduce + one D2H
            per_rank = max(n // utils.get_world_size(), 1)
            losses.update(m_loss, per_rank)
            top1.update(m_top1, per_rank)
            topk.update(m_topk, per_rank)
        batch_time.update(time.time() - end)
        end = time.time()
        if rank == 0 and ((idx + 1) % cfg.TRAIN.PRINT_FREQ == 0 or last):
            progress.cal_eta(idx + 1, n_iters, tic, cur_epoch, start_epoch)
            progress.display(idx + 1)
    watchdog.stop()
    heartbeat.beat(cur_epoch, n_iters, force=True)
    if cfg.B200.PROFILE and rank == 0:
        for name, st in timer.summary().items():
            per_gpu = st["n"] * targets.size(0) / max(st["total_ms"], 1e-9) * 1e3
            logger.info(f"PROFILE [{cur_epoch + 1}] {name}: {st['mean_ms']:.3f} ms/iter device time over {st['n']} iters "
                        f"({per_gpu:.1f} img/s per rank)")
    return losses.avg, top1.avg, topk.avg


def validate(val_loader, engine, device=None):
    """Evaluate; returns ``(top1.avg, topk.avg)`` (reference trainer.py:67-103)."""
    rank = utils.get_rank()
    device = device or next(engine.parameters()).device
    batch_time, data_time, losses, top1, topk = utils.construct_meters()
    n_iters = len(val_loader) if not cfg.B200.MAX_ITERS else min(len(val_loader), cfg.B200.MAX_ITERS)
    progress = utils.ProgressMeter(n_iters, [batch_time, data_time, losses, top1, topk], prefix="VAL:  ")
    if hasattr(engine, "sync_buffers"):
        engine.sync_buffers()   # every rank validates (and rank 0 checkpoints) the same running statistics
    engine.eval()
    metrics = utils.DeviceMetrics(device)
    sync_freq = max(int(cfg.B200.METRIC_SYNC_FREQ), 1)
    end = time.time()
    with torch.no_grad():
        for idx, (inputs, targets) in enumerate(val_loader):
            if idx >= n_iters:
                break
            data_time.update(time.time() - end)
            loss, hits1, hitsk = engine.eval_step(inputs, targets, cfg.TRAIN.TOPK)
            metrics.update(loss, hits1, hitsk, targets.size(0))
            last = (idx + 1) == n_iters
            if (idx + 1) % sync_freq == 0 or last:
                m_loss, m_top1, m_topk, n = metrics.flush()
                per_rank = max(n // utils.get_world_size(), 1)
                losses.update(m_loss, per_rank)
                top1.update(m_top1, per_rank)
                topk.update(m_topk, per_rank)
            batch_time.update(time.time() - end)
            end = time.time()
            if rank == 0 and ((idx + 1) % cfg.TEST.PRINT_FREQ == 0 or last):
                progress.display(idx + 1)
    return top1.avg, topk.avg


# ---------------------------------------------------------------------------------
# entry points
# ---------------------------------------------------------------------------------
def _bootstrap():
    utils.setup_distributed()
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ["LOCAL_RANK"])
    return rank, local_rank, utils.resolve_device()


def train_model():
    """Train according to the global ``cfg`` (reference trainer.py:106-173)."""
    rank, local_rank, device = _bootstrap()
    utils.setup_seed(rank)
    utils.setup_logger(rank, local_rank)

    net = build_net()
    native = _select_engine(device) == "native"
    if cfg.MODEL.SYNCBN and not native:  # the native engine fuses SyncBN itself
        net = SyncBatchNorm.convert_sync_batchnorm(net)
    net = net.to(device)
    engine = build_engine(net, device)

    train_loader = _wrap_loader(utils.construct_train_loader(device), device)
    val_loader = _wrap_loader(utils.construct_val_loader(device), device)
    optimizer = utils.construct_optimizer(engine)

    best_acc1 = start_epoch = 0
    if cfg.TRAIN.AUTO_RESUME and utils.has_checkpoint():
        start_epoch, best_acc1 = utils.load_checkpoint(utils.get_last_checkpoint(), engine, optimizer)
    elif cfg.MODEL.WEIGHTS:
        start_epoch, best_acc1 = utils.load_checkpoint(
            cfg.MODEL.WEIGHTS, engine, optimizer if cfg.TRAIN.LOAD_OPT else None)

    if rank == 0:
        logger.info("\n\n\n            =======  TRAINING  ======= \n\n")
        logger.info(utils.count_parameters(utils.unwrap_model(engine)))

    tic = time.time()
    acc1 = acck = 0.0
    for epoch in range(start_epoch, cfg.OPTIM.MAX_EPOCH):
        # Host-side rendezvous before any peer-memory kernel of the epoch is launched: the device-side waits of the
        # native engine are bounded (they trap instead of hanging the GPU), so skew from a slow checkpoint write or
        # a stalled loader has to be absorbed here, under the process group's DIST_TIMEOUT.
        utils.barrier()
        train_epoch(train_loader, engine, optimizer, epoch, start_epoch, tic, device)
        acc1, acck = validate(val_loader, engine, device)
        is_best = acc1 > best_acc1
        best_acc1 = max(acc1, best_acc1)
        checkpoint_file = utils.save_checkpoint(engine, optimizer, epoch, best_acc1, is_best)
        utils.barrier()         # nobody starts the next epoch's peer kernels while rank 0 is still writing
        if rank == 0:
            logger.info(f"ACCURACY: TOP1 {acc1:.3f}(BEST {best_acc1:.3f}) | "
                        f"TOP{cfg.TRAIN.TOPK} {acck:.3f} | SAVED {checkpoint_file}")
    utils.barrier()
    return best_acc1


def test_model():
    """Evaluate ``MODEL.WEIGHTS`` on the validation split (reference trainer.py:176-209)."""
    rank, local_rank, device = _bootstrap()
    utils.setup_logger(rank, local_rank)
    net = build_net().to(device)
    engine = build_engine(net, device)
    val_loader = _wrap_loader(utils.construct_val_loader(device), device)
    if cfg.MODEL.WEIGHTS:
        utils.load_checkpoint(cfg.MODEL.WEIGHTS, engine)
    acc1, acck = validate(val_loader, engine, device)
    if rank == 0:
        logger.info(f"ACCURACY: TOP1 {acc1:.3f}  |  TOP{cfg.TRAIN.TOPK} {acck:.3f}")
    return acc1, acck
