"""Cross-replica BatchNorm for the reference-semantics path (any device, any backend).

Parity: reference ``trainer.py:131`` converts every BatchNorm to
``torch.nn.SyncBatchNorm`` when ``MODEL.SYNCBN`` is set; that module only runs on CUDA
and issues one all_gather (fwd) + one all_reduce (bwd) per layer.  This
implementation reduces ``[sum, sum_sq, count]`` with a single all_reduce per direction, so
it also serves the CPU/gloo configuration; the native engine replaces it with the
peer-memory exchange inside the BN kernels (``csrc/elementwise.cu``: ``peer_exchange_reduce``).  State-dict keys equal ``nn.BatchNorm2d``'s.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class _SyncBNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, group):
        C = x.shape[1]
        dims = [0] + list(range(2, x.dim()))
        xf = x.float()
        stats = torch.empty(2 * C + 1, dtype=torch.float32, device=x.device)
        stats[:C] = xf.sum(dims)
        stats[C:2 * C] = (xf * xf).sum(dims)
        stats[2 * C] = x.numel() / C
        if _world(group) > 1:
            dist.all_reduce(stats, group=group)
        count = stats[2 * C]
        mean = stats[:C] / count
        var = (stats[C:2 * C] / count - mean * mean).clamp_min_(0.0)
        invstd = torch.rsqrt(var + eps)
        shape = [1, C] + [1] * (x.dim() - 2)
        xhat = (xf - mean.view(shape)) * invstd.view(shape)
        out = xhat * weight.float().view(shape) + bias.float().view(shape)
        ctx.save_for_backward(xhat, weight, invstd, count)
        ctx.group = group
        ctx.mark_non_differentiable(mean, var, count)
        return out.to(x.dtype), mean, var, count

    @staticmethod
    def backward(ctx, dout, _dm, _dv, _dc):
        xhat, weight, invstd, count = ctx.saved_tensors
        C = xhat.shape[1]
        dims = [0] + list(range(2, xhat.dim()))
        shape = [1, C] + [1] * (xhat.dim() - 2)
        g = dout.float()
        sums = torch.cat([g.sum(dims), (g * xhat).sum(dims)])
        dweight, dbias = sums[C:].clone(), sums[:C].clone()  # local sums: DDP averages them later
        if _world(ctx.group) > 1:
            dist.all_reduce(sums, group=ctx.group)
        mean_g = (sums[:C] / count).view(shape)
        mean_gx = (sums[C:] / count).view(shape)
        dx = (g - mean_g - xhat * mean_gx) * (weight.float() * invstd).view(shape)
        return dx.to(dout.dtype), dweight.to(weight.dtype), dbias.to(weight.dtype), None, None


class SyncBatchNorm(nn.modules.batchnorm._BatchNorm):
    """BatchNorm whose batch statistics span every rank of ``process_group``."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                 process_group=None):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)
        self.process_group = process_group

    def _check_input_dim(self, x):
        if x.dim() < 2:
            raise ValueError(f"expected at least 2D input (got {x.dim()}D input)")

    def forward(self, x):
        self._check_input_dim(x)
        use_batch_stats = self.training or not self.track_running_stats
        if not use_batch_stats or _world(self.process_group) == 1:
            return super().forward(x)
        w = self.weight if self.affine else torch.ones(self.num_features, device=x.device)
        b = self.bias if self.affine else torch.zeros(self.num_features, device=x.device)
        out, mean, var, count = _SyncBNFunction.apply(x, w, b, self.eps, self.process_group)
        if self.training and self.track_running_stats:
            with torch.no_grad():
                self.num_batches_tracked += 1
                mom = self.momentum if self.momentum is not None else 1.0 / float(self.num_batches_tracked)
                unbiased = var * (count / (count - 1).clamp_min(1.0))
                self.running_mean.mul_(1 - mom).add_(mean.to(self.running_mean.dtype), alpha=mom)
                self.running_var.mul_(1 - mom).add_(unbiased.to(self.running_var.dtype), alpha=mom)
        return out

    @classmethod
    def convert_sync_batchnorm(cls, module: nn.Module, process_group=None) -> nn.Module:
        """Recursively swap BatchNorm{1,2,3}d for ``SyncBatchNorm`` (parameters are shared)."""
        out = module
        if isinstance(module, nn.modules.batchnorm._BatchNorm) and not isinstance(module, cls):
            out = cls(module.num_features, module.eps, module.momentum, module.affine,
                      module.track_running_stats, process_group)
            if module.affine:
                out.weight, out.bias = module.weight, module.bias
            out.running_mean, out.running_var = module.running_mean, module.running_var
            out.num_batches_tracked = module.num_batches_tracked
            out.training = module.training
        for name, child in list(module.named_children()):
            new_child = cls.convert_sync_batchnorm(child, process_group)
            if new_child is not child:
                if isinstance(out, nn.Sequential) or isinstance(out, nn.ModuleDict) or isinstance(out, nn.ModuleList):
                    out._modules[name] = new_child
                else:
                    setattr(out, name, new_child)
        return out
