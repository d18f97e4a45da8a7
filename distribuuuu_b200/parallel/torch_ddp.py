"""Bucketed data parallelism on plain ``torch.distributed`` collectives.

This is the *reference-semantics* engine: it does what ``DistributedDataParallel`` does
for the reference (trainer.py:134) -- broadcast the initial state from rank
0, average gradients over ranks in size-capped buckets launched as soon as their last
gradient is produced, overlapping with the rest of backward -- but through public
collectives only, so it runs on gloo/CPU as well as NCCL.  The native engine
(``parallel.native_engine``) replaces the bucket all-reduce + ``optimizer.step`` with one
fused peer-memory kernel; this class is its fallback and its numerical reference.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist
import torch.nn as nn

_FIRST_BUCKET_BYTES = 1 << 20  # same as torch DDP's first bucket


class _Bucket:
    __slots__ = ("params", "flat", "views", "pending", "work", "offsets")

    def __init__(self, params: List[nn.Parameter]):
        self.params = params
        total = sum(p.numel() for p in params)
        ref = params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self.views, self.offsets, off = [], [], 0
        for p in params:
            self.views.append(self.flat[off:off + p.numel()].view(p.shape))
            self.offsets.append(off)
            off += p.numel()
        self.pending = len(params)
        self.work = None


def plan_buckets(params: List[nn.Parameter], cap_bytes: int, first_cap_bytes: int = _FIRST_BUCKET_BYTES):
    """Greedy bucket assignment over parameters in *reverse* registration order (the order
    gradients become ready); the first bucket is small so communication starts early."""
    buckets, cur, cur_bytes = [], [], 0
    cap = first_cap_bytes
    for p in reversed(params):
        nbytes = p.numel() * p.element_size()
        if cur and (cur_bytes + nbytes > cap or p.dtype != cur[0].dtype):
            buckets.append(cur)
            cur, cur_bytes, cap = [], 0, cap_bytes
        cur.append(p)
        cur_bytes += nbytes
    if cur:
        buckets.append(cur)
    return buckets


class BucketedDataParallel(nn.Module):
    """Wraps ``module``; after ``loss.backward()`` call ``finish_backward()`` (done by the
    trainer, or implicitly by the optimizer hook installed here) before ``optimizer.step``."""

    def __init__(self, module: nn.Module, bucket_cap_mb: float = 25, process_group=None,
                 broadcast_buffers: bool = True):
        super().__init__()
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.broadcast_buffers = broadcast_buffers
        self._params = [p for p in module.parameters() if p.requires_grad]
        self._sync_initial_state()
        self._buckets: List[_Bucket] = []
        self._where = {}
        self._hooks = []
        self._launched = 0
        if self.world > 1:
            for plist in plan_buckets(self._params, int(bucket_cap_mb * 1024 * 1024)):
                b = _Bucket(plist)
                for i, p in enumerate(plist):
                    self._where[p] = (b, i)
                self._buckets.append(b)
            for p in self._params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad_ready))
        self._need_buffer_sync = False

    # -- state synchronisation ------------------------------------------------------
    def _sync_initial_state(self):
        if self.world == 1:
            return
        with torch.no_grad():
            for t in list(self.module.parameters()) + list(self.module.buffers()):
                dist.broadcast(t.data, src=0, group=self.group)

    def _sync_buffers(self):
        """Rank 0's buffers (BN running stats) win, like DDP's broadcast_buffers=True."""
        if self.world == 1 or not self.broadcast_buffers:
            return
        bufs = [b for b in self.module.buffers() if b.is_floating_point()]
        if not bufs:
            return
        with torch.no_grad():
            flat = torch.cat([b.reshape(-1).float() for b in bufs])
            dist.broadcast(flat, src=0, group=self.group)
            off = 0
            for b in bufs:
                b.copy_(flat[off:off + b.numel()].view(b.shape))
                off += b.numel()

    def sync_buffers(self):
        """Make every rank evaluate / checkpoint with rank 0's buffers.  Reference DDP does this implicitly on the
        first eval forward after training (``require_forward_param_sync`` is still set, trainer.py:134), so the
        accuracy it logs is reproducible from the checkpoint rank 0 writes."""
        self._sync_buffers()
        self._need_buffer_sync = False

    # -- gradient path ------------------------------------------------------------------
    def _on_grad_ready(self, p: nn.Parameter):
        b, i = self._where[p]
        b.views[i].copy_(p.grad)
        p.grad = b.views[i]  # gradient-as-bucket-view: no copy back after the reduce
        b.pending -= 1
        if b.pending == 0:
            b.flat.div_(self.world)
            b.work = dist.all_reduce(b.flat, group=self.group, async_op=True)
            self._launched += 1

    def finish_backward(self):
        """Wait for every bucket; unused parameters get a zero gradient contribution."""
        if self.world == 1:
            return
        for b in self._buckets:
            if b.work is None:
                # some parameter had no gradient this step: contribute zeros for it
                for i, p in enumerate(b.params):
                    if p.grad is None or p.grad.data_ptr() != b.views[i].data_ptr():
                        if p.grad is None:
                            b.views[i].zero_()
                        else:
                            b.views[i].copy_(p.grad)
                        p.grad = b.views[i]
                b.flat.div_(self.world)
                b.work = dist.all_reduce(b.flat, group=self.group, async_op=True)
        for b in self._buckets:
            b.work.wait()
            b.work = None
            b.pending = len(b.params)
        self._launched = 0

    def forward(self, *args, **kwargs):
        if self.training and torch.is_grad_enabled() and self._need_buffer_sync:
            self._sync_buffers()
        out = self.module(*args, **kwargs)
        self._need_buffer_sync = self.training and torch.is_grad_enabled()
        return out

    def zero_grad(self, set_to_none: bool = True):
        # bucket views are reused: zeroing is unnecessary because hooks overwrite them
        for p in self._params:
            if p.grad is not None and p in self._where:
                b, i = self._where[p]
                if p.grad.data_ptr() == b.views[i].data_ptr():
                    p.grad = None
                    continue
            p.grad = None
