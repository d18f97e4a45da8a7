"""Shared pieces of the tutorial ladder (CIFAR-10 / ResNet-18): data, model, train loop.

The ladder mirrors the reference's ``tutorial/`` scripts (snsc -> snmc_dp -> mnmc_ddp_launch ->
mnmc_ddp_mp -> mnmc_ddp_slurm -> imagenet) but every script works on CPU or GPU and offline
(``--synthetic`` is implied when the dataset cannot be found, since this image has no network).
"""
from __future__ import annotations

import argparse
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distribuuuu_b200 import models  # noqa: E402


def base_parser(desc: str) -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description=desc)
    ap.add_argument("--data", default="./data", help="CIFAR-10 root (torchvision layout)")
    ap.add_argument("--synthetic", action="store_true", help="random CIFAR-shaped data")
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--batch-size", type=int, default=128, help="per process")
    ap.add_argument("--lr", type=float, default=0.01, help="per-process base LR (scaled by world size in DDP scripts)")
    ap.add_argument("--workers", type=int, default=2)
    ap.add_argument("--print-freq", type=int, default=25)
    ap.add_argument("--max-iters", type=int, default=0, help="stop each epoch early (smoke runs)")
    ap.add_argument("--device", default="auto", choices=["auto", "cuda", "cpu"])
    return ap


def pick_device(choice: str, local_rank: int = 0) -> torch.device:
    if choice == "cpu" or (choice == "auto" and not torch.cuda.is_available()):
        return torch.device("cpu")
    return torch.device("cuda", local_rank % torch.cuda.device_count())


class SyntheticCifar(torch.utils.data.Dataset):
    def __init__(self, n=2048, classes=10, size=32):
        g = torch.Generator().manual_seed(0)
        self.y = torch.randint(0, classes, (n,), generator=g)
        # class-dependent mean so the task is learnable and accuracy moves
        self.x = torch.randn(n, 3, size, size, generator=g) + self.y.view(-1, 1, 1, 1).float() / classes

    def __len__(self):
        return len(self.y)

    def __getitem__(self, i):
        return self.x[i], int(self.y[i])


def cifar10(root: str, synthetic: bool):
    if not synthetic:
        try:
            import torchvision
            import torchvision.transforms as T
            tf = T.Compose([T.RandomCrop(32, padding=4), T.RandomHorizontalFlip(), T.ToTensor(),
                            T.Normalize((0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010))])
            return torchvision.datasets.CIFAR10(root=root, train=True, download=False, transform=tf)
        except Exception as exc:  # no dataset on disk and no network
            print(f"[tutorial] CIFAR-10 not found under {root} ({type(exc).__name__}); using synthetic data")
    return SyntheticCifar()


def make_net(num_classes: int = 10) -> nn.Module:
    return models.resnet18(num_classes=num_classes)


def train_one_epoch(net, loader, optimizer, device, epoch, rank=0, print_freq=25, max_iters=0, sampler=None):
    if sampler is not None:
        sampler.set_epoch(epoch)  # reshuffle differently every epoch, identically on every rank
    criterion = nn.CrossEntropyLoss()
    net.train()
    loss_sum = correct = seen = 0
    for idx, (x, y) in enumerate(loader):
        if max_iters and idx >= max_iters:
            break
        x, y = x.to(device, non_blocking=True), y.to(device, non_blocking=True)
        out = net(x)
        loss = criterion(out, y)
        optimizer.zero_grad(set_to_none=True)
        loss.backward()
        optimizer.step()
        loss_sum += loss.item() * y.size(0)
        correct += (out.argmax(1) == y).sum().item()
        seen += y.size(0)
        if rank == 0 and ((idx + 1) % print_freq == 0 or (idx + 1) == len(loader)):
            print(f"   == step: [{idx + 1:3d}/{len(loader)}] [{epoch}] | loss: {loss_sum / seen:.3f} | "
                  f"acc: {100.0 * correct / seen:6.3f}%", flush=True)
    return loss_sum / max(seen, 1), 100.0 * correct / max(seen, 1)
