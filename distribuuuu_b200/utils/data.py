"""Datasets and loaders.

Parity: reference ``utils.py:109-184``: ``DummyDataset`` (host randn, label 0),
ImageFolder train pipeline (RandomResizedCrop(IM_SIZE) / HFlip / ToTensor / Normalize,
DistributedSampler(shuffle), drop_last) and val pipeline (Resize(TEST.IM_SIZE) /
CenterCrop(224) / ..., padded DistributedSampler, keep last).

New: ``B200.INPUT_UINT8`` keeps images as uint8 through the loader and the H2D copy (4x fewer bytes) and lets the
GPU normalise them (native engine: inside the stem im2col kernel); ``SyntheticDeviceLoader`` generates ImageNet-shaped batches directly on the GPU
(no 602 MB host tensor, no zero-iteration corner when ranks x batch > 1000)
and ``PinnedPrefetcher`` overlaps the H2D copy of batch i+1 with step i.
"""
from __future__ import annotations

import os

import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler

from ..config import cfg
from .dist import get_rank, get_world_size

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def normalize_uint8(x: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    """``(x / 255 - mean) / std`` for a uint8 NCHW batch (what ``ToTensor`` + ``Normalize`` do on the host in the
    reference, utils.py:127-135).  The native engine does this inside its stem kernel instead; this is the
    ATen version used by the torch engine and by tests."""
    m = torch.tensor(mean, dtype=torch.float32, device=x.device).view(1, -1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32, device=x.device).view(1, -1, 1, 1)
    return (x.float() / 255.0 - m) / s


class DummyDataset(Dataset):
    """``length`` random images of shape ``size``; every label is 0.  ``uint8=True`` yields raw-pixel images
    (for ``B200.INPUT_UINT8``)."""

    def __init__(self, length: int, size, uint8: bool = False):
        self.len = int(length)
        gen = torch.Generator().manual_seed(0)
        if uint8:
            self.data = torch.randint(0, 256, [self.len] + list(size), generator=gen, dtype=torch.uint8)
        else:
            self.data = torch.randn([self.len] + list(size), generator=gen)

    def __getitem__(self, index):
        return self.data[index], 0

    def __len__(self):
        return self.len


class _NoopSampler:
    def set_epoch(self, epoch: int) -> None:  # same surface as DistributedSampler
        self.epoch = epoch


class SyntheticDeviceLoader:
    """Endless-epoch-free synthetic source living on the compute device.

    Yields ``iters`` batches per epoch of ``(randn[B,3,S,S], randint[B])``; images are
    regenerated per batch by a device RNG so nothing crosses PCIe.
    """

    def __init__(self, batch_size: int, im_size: int, num_classes: int, length: int,
                 device: torch.device, drop_last: bool = True, dtype=torch.float32):
        per_rank = length // get_world_size() if drop_last else -(-length // get_world_size())
        self.iters = per_rank // batch_size if drop_last else -(-per_rank // batch_size)
        self.batch_size, self.im_size, self.num_classes = batch_size, im_size, num_classes
        self.device, self.dtype = device, dtype
        self.sampler = _NoopSampler()
        self.gen = torch.Generator(device=device).manual_seed(1234 + get_rank())

    def __len__(self):
        return self.iters

    def __iter__(self):
        for _ in range(self.iters):
            if self.dtype == torch.uint8:
                x = torch.randint(0, 256, (self.batch_size, 3, self.im_size, self.im_size), device=self.device,
                                  dtype=torch.uint8, generator=self.gen)
            else:
                x = torch.randn(self.batch_size, 3, self.im_size, self.im_size, device=self.device,
                                dtype=self.dtype, generator=self.gen)
            y = torch.randint(0, self.num_classes, (self.batch_size,), device=self.device, generator=self.gen)
            yield x, y


_SIDE_STREAMS: dict = {}


def _side_stream(device: torch.device):
    """One copy stream per device for the whole process (a fresh stream per loader would also mean a fresh, cold
    caching-allocator pool per epoch)."""
    if device.type != "cuda":
        return None
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


class PinnedPrefetcher:
    """Wraps a host loader: copies batch i+1 to the device on a side stream while the
    caller computes on batch i (replaces the blocking ``inputs.cuda()`` at reference
    trainer.py:40).

    Batches land in two persistent device staging slots owned by the prefetcher, so the steady state performs no
    allocation at all: slot k is overwritten (on the copy stream) only after an event recorded on the compute
    stream once the step that consumed it has been enqueued.  The yielded tensors are views of those slots and
    are valid until the batch after next is requested -- i.e. for the whole training/eval step."""

    SLOTS = 2

    def __init__(self, loader, device: torch.device):
        self.loader, self.device = loader, device
        self.sampler = getattr(loader, "sampler", _NoopSampler())
        self.stream = _side_stream(device)
        self._bufs: dict = {}
        self._free = [None] * self.SLOTS

    def __len__(self):
        return len(self.loader)

    def _slot(self, k: int, name: str, like: torch.Tensor) -> torch.Tensor:
        buf = self._bufs.get((k, name))
        if (buf is None or buf.dtype != like.dtype or buf.shape[1:] != like.shape[1:] or buf.shape[0] < like.shape[0]):
            buf = torch.empty(like.shape, dtype=like.dtype, device=self.device)   # compute-stream pool, allocated once
            self._bufs[(k, name)] = buf
        return buf[: like.shape[0]]

    def _stage(self, batch, k: int):
        x, y = batch
        if not torch.is_tensor(y):
            y = torch.as_tensor(y)
        if self.stream is None:
            return x.to(self.device), y.to(self.device), None
        bx, by = self._slot(k, "x", x), self._slot(k, "y", y)
        with torch.cuda.stream(self.stream):
            if self._free[k] is not None:
                self.stream.wait_event(self._free[k])       # the step that read this slot has been enqueued and finished
            else:
                self.stream.wait_stream(torch.cuda.current_stream(self.device))  # first use: order after the allocation
            bx.copy_(x, non_blocking=True)
            by.copy_(y, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        return bx, by, done

    def __iter__(self):
        it = iter(self.loader)
        k = 0
        try:
            nxt = self._stage(next(it), k)
        except StopIteration:
            return
        try:
            while nxt is not None:
                x, y, done = nxt
                cur_k = k
                if done is not None:
                    torch.cuda.current_stream(self.device).wait_event(done)
                k = (k + 1) % self.SLOTS
                try:
                    nxt = self._stage(next(it), k)
                except StopIteration:
                    nxt = None
                yield x, y
                if self.stream is not None:   # the consumer has enqueued its step: after it, slot cur_k may be overwritten
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(self.device))
                    self._free[cur_k] = ev
        finally:
            # consumer stopped early: order the compute stream after the copy that is still in flight, so the slot
            # can never be recycled underneath it
            if nxt is not None and nxt[2] is not None:
                torch.cuda.current_stream(self.device).wait_event(nxt[2])


def _transforms():
    import torchvision.transforms as T
    return T


def _image_folder(root, transform):
    import torchvision
    return torchvision.datasets.ImageFolder(root=root, transform=transform)


def _use_device_synthetic() -> bool:
    return bool(cfg.MODEL.DUMMY_INPUT and cfg.B200.DUMMY_ON_DEVICE)


def _tail_transforms(T):
    """Last stage of the image pipeline: the reference's ToTensor + Normalize (fp32, 12 bytes/pixel), or -- with
    ``B200.INPUT_UINT8`` -- the raw uint8 CHW tensor (3 bytes/pixel over PCIe; normalised on the GPU)."""
    if cfg.B200.INPUT_UINT8:
        return [T.PILToTensor()]
    return [T.ToTensor(), T.Normalize(mean=IMAGENET_MEAN, std=IMAGENET_STD)]


def _input_dtype():
    return torch.uint8 if cfg.B200.INPUT_UINT8 else torch.float32


def construct_train_loader(device: torch.device | None = None):
    if _use_device_synthetic() and device is not None:
        return SyntheticDeviceLoader(cfg.TRAIN.BATCH_SIZE, cfg.TRAIN.IM_SIZE, cfg.MODEL.NUM_CLASSES,
                                     cfg.B200.DUMMY_LEN, device, drop_last=True, dtype=_input_dtype())
    if cfg.MODEL.DUMMY_INPUT:
        trainset = DummyDataset(cfg.B200.DUMMY_LEN, [3, cfg.TRAIN.IM_SIZE, cfg.TRAIN.IM_SIZE], uint8=cfg.B200.INPUT_UINT8)
    else:
        T = _transforms()
        trainset = _image_folder(os.path.join(cfg.TRAIN.DATASET, cfg.TRAIN.SPLIT), T.Compose([
            T.RandomResizedCrop(cfg.TRAIN.IM_SIZE), T.RandomHorizontalFlip()] + _tail_transforms(T)))
    sampler = DistributedSampler(trainset, num_replicas=get_world_size(), rank=get_rank(), shuffle=True)
    return DataLoader(trainset, batch_size=cfg.TRAIN.BATCH_SIZE, num_workers=cfg.TRAIN.WORKERS,
                      pin_memory=cfg.TRAIN.PIN_MEMORY and torch.cuda.is_available(),
                      sampler=sampler, drop_last=True, persistent_workers=cfg.TRAIN.WORKERS > 0)


def construct_val_loader(device: torch.device | None = None):
    if _use_device_synthetic() and device is not None:
        return SyntheticDeviceLoader(cfg.TEST.BATCH_SIZE, 224, cfg.MODEL.NUM_CLASSES,
                                     cfg.B200.DUMMY_LEN, device, drop_last=False, dtype=_input_dtype())
    if cfg.MODEL.DUMMY_INPUT:
        valset = DummyDataset(cfg.B200.DUMMY_LEN, [3, 224, 224], uint8=cfg.B200.INPUT_UINT8)
    else:
        T = _transforms()
        # the reference reads the val split under TRAIN.DATASET (utils.py:157); kept.
        valset = _image_folder(os.path.join(cfg.TRAIN.DATASET, cfg.TEST.SPLIT), T.Compose([
            T.Resize(cfg.TEST.IM_SIZE), T.CenterCrop(224)] + _tail_transforms(T)))
    sampler = DistributedSampler(valset, num_replicas=get_world_size(), rank=get_rank(), shuffle=False)
    return DataLoader(valset, batch_size=cfg.TEST.BATCH_SIZE, shuffle=False, sampler=sampler,
                      num_workers=cfg.TRAIN.WORKERS, pin_memory=cfg.TRAIN.PIN_MEMORY and torch.cuda.is_available(),
                      drop_last=False, persistent_workers=cfg.TRAIN.WORKERS > 0)
