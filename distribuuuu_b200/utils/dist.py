"""Process bootstrap and small collectives.

Parity: reference ``distribuuuu/utils.py:19-51`` (``setup_distributed``: Slurm branch
derives RANK/WORLD_SIZE/LOCAL_RANK/MASTER_ADDR from the scheduler, launch branch reads
torchrun's env) and ``utils.py:85-106`` (``scaled_all_reduce``).

Differences by design: backend and device are selectable (nccl+cuda on a GPU box,
gloo+cpu anywhere) so the same code serves the CPU plumbing config of BASELINE.json,
and the three per-iteration metric reductions are packed into one message.
"""
from __future__ import annotations

import os
import subprocess
from datetime import timedelta

import torch
import torch.distributed as dist

from ..config import cfg

_SLURM_DEFAULT_PORT = "29566"  # reference utils.py:35


def resolve_device() -> torch.device:
    """Device for this rank according to ``cfg.B200.DEVICE`` and LOCAL_RANK."""
    want = cfg.B200.DEVICE
    if want == "auto":
        want = "cuda" if torch.cuda.is_available() else "cpu"
    if want == "cuda":
        if not torch.cuda.is_available():
            raise RuntimeError("B200.DEVICE=cuda but no CUDA device is visible")
        local_rank = int(os.environ.get("LOCAL_RANK", 0))
        return torch.device("cuda", local_rank % max(torch.cuda.device_count(), 1))
    return torch.device("cpu")


def resolve_backend(device: torch.device, backend: str | None = None) -> str:
    backend = backend or cfg.B200.DIST_BACKEND
    if backend in (None, "auto"):
        backend = "nccl" if device.type == "cuda" else "gloo"
    return backend


def _first_slurm_host(node_list: str) -> str:
    out = subprocess.getoutput(f"scontrol show hostname {node_list} | head -n1").strip()
    return out or "127.0.0.1"


def setup_distributed(backend: str | None = None, port: int | str | None = None) -> None:
    """Initialise ``torch.distributed`` from Slurm or launcher environment variables.

    Slurm (``SLURM_JOB_ID`` present): rank = ``SLURM_PROCID``, world = ``SLURM_NTASKS``,
    master = first host of ``SLURM_NODELIST``; the derived values are exported so that
    the rest of the code (and child processes) see the torchrun contract.
    Otherwise ``RANK``/``WORLD_SIZE`` (+ ``LOCAL_RANK``, ``MASTER_ADDR``, ``MASTER_PORT``)
    must be set by ``torchrun`` / ``torch.distributed.launch --use_env`` / ``mp.spawn``.
    A plain ``python train_net.py`` run (no env at all) becomes a world of one.
    """
    if dist.is_available() and dist.is_initialized():
        return
    n_local = torch.cuda.device_count() if torch.cuda.is_available() else 0
    env = os.environ
    if "SLURM_JOB_ID" in env and "SLURM_PROCID" in env:
        rank = int(env["SLURM_PROCID"])
        world = int(env["SLURM_NTASKS"])
        if port is not None:
            env["MASTER_PORT"] = str(port)
        env.setdefault("MASTER_PORT", _SLURM_DEFAULT_PORT)
        if "MASTER_ADDR" not in env:
            env["MASTER_ADDR"] = _first_slurm_host(env.get("SLURM_NODELIST", ""))
        env["WORLD_SIZE"] = str(world)
        env["RANK"] = str(rank)
        env["LOCAL_RANK"] = str(int(env.get("SLURM_LOCALID", rank % max(n_local, 1))))
    else:
        rank = int(env.setdefault("RANK", "0"))
        world = int(env.setdefault("WORLD_SIZE", "1"))
        env.setdefault("LOCAL_RANK", str(rank % max(n_local, 1)))
        env.setdefault("MASTER_ADDR", "127.0.0.1")
        env.setdefault("MASTER_PORT", str(port) if port is not None else "29500")

    device = resolve_device()
    if device.type == "cuda":
        torch.cuda.set_device(device)
    backend = resolve_backend(device, backend)
    kwargs = dict(backend=backend, world_size=world, rank=rank, timeout=timedelta(minutes=float(cfg.B200.DIST_TIMEOUT_MIN)))
    if device.type == "cuda":
        kwargs["device_id"] = device  # eager NCCL communicator, needed for symmetric memory
    dist.init_process_group(**kwargs)


def get_rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def get_world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def is_primary() -> bool:
    return get_rank() == 0


def barrier() -> None:
    if get_world_size() > 1:
        dist.barrier()


def scaled_all_reduce(tensors):
    """Sum-reduce every tensor over the job and scale by 1/world, in place.

    Semantics of reference utils.py:85-106; the tensors are packed so one collective is
    issued instead of ``len(tensors)``.
    """
    world = get_world_size()
    if world == 1:
        return tensors
    flat = torch.cat([t.detach().reshape(-1).float() for t in tensors])
    dist.all_reduce(flat)
    flat.mul_(1.0 / world)
    off = 0
    for t in tensors:
        n = t.numel()
        t.detach().copy_(flat[off:off + n].view(t.shape))
        off += n
    return tensors


def broadcast_object(obj, src: int = 0):
    if get_world_size() == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def shutdown() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
