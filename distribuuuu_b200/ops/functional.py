"""Functional building blocks every model in the zoo is written against.

Each function takes the *module(s)* holding the parameters (so ``state_dict`` keys stay
torchvision/timm compatible) and dispatches on ``runtime.active_engine()``:

* no engine  -> plain torch ops (reference semantics: conv -> BN -> act as separate
  ATen/cuDNN calls, exactly what reference resnet.py:84-103 / 142-161 executes);
* engine     -> fused sm_100a kernels from ``ops.native`` (conv implicit-GEMM with the
  BN statistics in its epilogue, one normalise+affine+act(+residual) pass, peer-memory
  SyncBN, ...).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import runtime


def _act(x, act):
    if act is None:
        return x
    if act == "relu":
        return F.relu(x, inplace=True)
    if act == "silu":
        return F.silu(x, inplace=True)
    if act == "sigmoid":
        return torch.sigmoid(x)
    raise ValueError(f"unknown activation {act!r}")


def conv_bn_act(x, conv: nn.Conv2d, bn: nn.Module | None, act: str | None = None, residual=None, residual_sink=None,
                input_grad_to=None, out_buffer=None):
    """``act(bn(conv(x)) + residual)``; ``bn``/``act``/``residual`` optional.

    ``residual_sink``: the conv module (of the same block) that consumes the *same tensor* as ``residual``; a hint
    that lets the native engine fold the residual-branch gradient into that conv's dgrad epilogue.
    ``input_grad_to``: a conv module applied EARLIER in this forward to the *same tensor* ``x`` (a projection
    shortcut); this conv's input gradient is then handed to that conv's dgrad instead of being summed by a separate
    kernel.  Both are hints: ignored on the torch path and whenever the tensors do not actually coincide.
    ``out_buffer``: handle from ``dense_block_buffer``; a plain conv then stores its output straight into the block's
    feature buffer (no later concatenation copy)."""
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.conv_bn_act(x, conv, bn, act, residual, residual_sink, input_grad_to, out_buffer)
    y = conv(x)
    if bn is not None:
        y = bn(y)
    if residual is not None:
        y = y + residual
    return _act(y, act)


def conv_bn_relu_maxpool(x, conv: nn.Conv2d, bn: nn.Module, kernel_size: int = 3, stride: int = 2, padding: int = 1):
    """The classifier stem ``maxpool(relu(bn(conv(x))))`` (reference models/resnet.py:194-197, densenet.py, botnet.py).
    The native engine normalises, rectifies and pools in one pass over the conv output and back-propagates straight from
    the pooled gradient, so the full-resolution activation is never stored."""
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.conv_bn_relu_maxpool(x, conv, bn, kernel_size, stride, padding)
    return F.max_pool2d(F.relu(bn(conv(x)), inplace=True), kernel_size, stride, padding)


def bn_act(x, bn: nn.Module, act: str | None = None):
    """Pre-activation normalisation (DenseNet ordering: BN -> ReLU -> conv)."""
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.bn_act(x, bn, act)
    return _act(bn(x), act)


def conv2d(x, conv: nn.Conv2d, act: str | None = None, out_buffer=None):
    return conv_bn_act(x, conv, None, act, out_buffer=out_buffer)


def dense_block_buffer(x, total_channels: int):
    """DenseNet blocks: returns ``(x, handle)``.  With the native engine, ``x`` becomes a view into one pre-allocated
    NHWC buffer of ``total_channels`` channels and ``handle`` lets every layer's conv write its new features right
    behind the previous ones, so ``concat_channels`` of the block's features is a view (reference densenet.py:68,148
    copies them with ``torch.cat`` in every layer).  On the torch path: ``(x, None)``."""
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.dense_block_buffer(x, total_channels)
    return x, None


def linear(x, fc: nn.Linear):
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.linear(x, fc)
    return fc(x)


def max_pool2d(x, kernel_size: int, stride: int, padding: int = 0):
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.max_pool2d(x, kernel_size, stride, padding)
    return F.max_pool2d(x, kernel_size, stride, padding)


def avg_pool2d(x, kernel_size: int, stride: int | None = None):
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.avg_pool2d(x, kernel_size, stride or kernel_size)
    return F.avg_pool2d(x, kernel_size, stride or kernel_size)


def global_avg_pool(x):
    """[N,C,H,W] -> [N,C]."""
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.global_avg_pool(x)
    return x.mean(dim=(2, 3))


def squeeze_excite(x, fc1: nn.Conv2d, fc2: nn.Conv2d, act: str = "relu"):
    """x * sigmoid(fc2(act(fc1(mean_hw(x))))) -- SE gate used by RegNetY / EfficientNet."""
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.squeeze_excite(x, fc1, fc2, act)
    s = x.mean(dim=(2, 3), keepdim=True)
    s = _act(fc1(s), act)
    return x * torch.sigmoid(fc2(s))


def concat_channels(tensors):
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.concat_channels(tensors)
    return torch.cat(list(tensors), dim=1)


def dropout(x, p: float, training: bool):
    if p <= 0.0 or not training:
        return x
    return F.dropout(x, p=p, training=True)


def relpos_attention(q, k, v, rel_h, rel_w, height: int, width: int, scale: float):
    """Multi-head self-attention over an HxW map with 2-D relative-position logits.

    q,k: [B, heads, H*W, dqk]; v: [B, heads, H*W, dv]; rel_h: [2H-1, dqk]; rel_w: [2W-1, dqk].
    ``logit[(x,y),(i,j)] = s*q.k + s*q.rel_w[j-y+W-1] + s*q.rel_h[i-x+H-1]`` (reference
    botnet.py:25-57,77-98,206-212 computes the same thing through pad/reshape tricks).
    """
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.relpos_attention(q, k, v, rel_h, rel_w, height, width, scale)
    B, nh, L, d = q.shape
    q = q * scale
    logits = torch.matmul(q, k.transpose(-1, -2))
    q5 = q.reshape(B, nh, height, width, d)
    dev = q.device
    iw = torch.arange(width, device=dev)
    ih = torch.arange(height, device=dev)
    tab_w = rel_w[(iw[None, :] - iw[:, None]) + width - 1].to(q.dtype)   # [y, j, d]
    tab_h = rel_h[(ih[None, :] - ih[:, None]) + height - 1].to(q.dtype)  # [x, i, d]
    lw = torch.einsum("bnxyd,yjd->bnxyj", q5, tab_w)
    lh = torch.einsum("bnxyd,xid->bnxyi", q5, tab_h)
    pos = lh[:, :, :, :, :, None] + lw[:, :, :, :, None, :]              # [b,n,x,y,i,j]
    logits = logits + pos.reshape(B, nh, L, L)
    return torch.matmul(torch.softmax(logits, dim=-1), v)


def relpos_mhsa(qk, v, rel_h, rel_w, heads: int, dim_qk: int, dim_v: int, scale: float):
    """BoTNet's MHSA core on the raw projections: ``qk`` [B, 2*heads*dim_qk, H, W] (q heads first, then k heads),
    ``v`` [B, heads*dim_v, H, W] -> [B, heads*dim_v, H, W].  Same math as splitting into heads and calling
    ``relpos_attention`` (reference botnet.py:193-215); the native engine runs it as ONE fused tcgen05 kernel on the
    NHWC projections (no head split / transpose copies) when the map is 14x14 with 128-wide heads."""
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.relpos_mhsa(qk, v, rel_h, rel_w, heads, dim_qk, dim_v, scale)
    return _relpos_mhsa_composite(qk, v, rel_h, rel_w, heads, dim_qk, dim_v, scale)


def _relpos_mhsa_composite(qk, v, rel_h, rel_w, heads, dim_qk, dim_v, scale):
    B, _, H, W = qk.shape
    q, k = qk[:, : heads * dim_qk], qk[:, heads * dim_qk:]

    def split(t, d):  # [B,(h d),H,W] -> [B,h,HW,d]
        return t.reshape(B, heads, d, H * W).transpose(2, 3)

    out = relpos_attention(split(q, dim_qk), split(k, dim_qk), split(v, dim_v), rel_h, rel_w, H, W, scale)
    return out.transpose(2, 3).reshape(B, heads * dim_v, H, W)


def abspos_attention(q, k, v, emb_h, emb_w, scale: float):
    """Attention with absolute position logits q.(emb_h[x]+emb_w[y]) (reference botnet.py:60-74)."""
    q = q * scale
    emb = (emb_h[:, None, :] + emb_w[None, :, :]).reshape(-1, emb_h.shape[-1]).to(q.dtype)
    logits = torch.matmul(q, k.transpose(-1, -2)) + torch.matmul(q, emb.t())
    return torch.matmul(torch.softmax(logits, dim=-1), v)


def cross_entropy_topk(logits, target, topk: int = 5):
    """Mean CE loss plus top-1 / top-k hit *counts* (0-dim tensors) in one call
    (reference trainer.py:43,50: ``CrossEntropyLoss`` then ``utils.accuracy``)."""
    eng = runtime.active_engine()
    if eng is not None:
        return eng.ops.cross_entropy_topk(logits, target, topk)
    loss = F.cross_entropy(logits.float(), target)
    with torch.no_grad():
        k = min(topk, logits.size(1))
        top = logits.topk(k, dim=1).indices
        hit = top.eq(target.view(-1, 1))
        return loss, hit[:, :1].sum(), hit.any(dim=1).sum()
