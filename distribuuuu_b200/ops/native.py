"""Autograd wrappers around the sm_100a kernels (``csrc/``) -- the *native* implementation of
``ops.functional``.

Conventions on this path: activations are logical ``[N, C, H, W]`` bf16 tensors in
``channels_last`` memory format (physically NHWC, which is what the kernels index); weights
come from the engine's flat bf16 buffer as contiguous ``[Cout, R, S, Cin]`` views; weight
gradients are written by the kernels *directly* into the engine's flat fp32 gradient buffer
(no ``.grad`` tensors, no bucket copies -- SURVEY G19) and the engine is told when a parameter
is ready so the fused all-reduce + SGD kernel can start while backward is still running.

What still runs as plain torch ops on the same bf16 NHWC tensors (their parameter gradients are routed into the
same flat buffer by a hook, see ``NativeEngine.w16_leaf``): convs / BN whose channel count is not a multiple of 8
and convolutions with a bias or asymmetric geometry -- none of which occurs in the zoo's ``config/*.yaml`` models
(``NativeOps.fallbacks`` counts every such call; tests assert it stays empty for them).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

ACT = {None: 0, "relu": 1, "silu": 2}


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """Contiguous [N,H,W,C] view (copy only if the tensor is not channels_last already)."""
    v = x.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


def _nchw_view(x_nhwc: torch.Tensor) -> torch.Tensor:
    return x_nhwc.permute(0, 3, 1, 2)


def _rows_view(x: torch.Tensor):
    """``([rows, C] view with a row pitch, (N, H, W, C))`` of a logical-NCHW activation WITHOUT copying when it is NHWC
    with a uniform pixel pitch -- contiguous tensors and channel slices of a wider NHWC buffer (DenseNet blocks)."""
    v = x.permute(0, 2, 3, 1)
    N, H, W, C = v.shape
    pitch = v.stride(2)
    if v.stride(3) == 1 and pitch >= C and v.stride(1) == W * pitch and v.stride(0) == H * W * pitch:
        return torch.as_strided(v, (N * H * W, C), (pitch, 1), v.storage_offset()), (N, H, W, C)
    v = v.contiguous()
    return v.view(-1, C), (N, H, W, C)


class _DenseBuffer:
    """One pre-allocated NHWC buffer per dense block: the block input is placed in its first channels and every layer's
    3x3 convolution TMA-stores its ``growth_rate`` new channels right behind -- the concatenated input of layer l is
    the first c_l channels of the buffer, a view (reference densenet.py:68,148 materialises it with torch.cat)."""

    def __init__(self, buf: torch.Tensor, used: int):
        self.buf, self.used = buf, used

    def take(self, channels: int):
        """Channel slice [used, used + channels) as a logical-NCHW view, or None when the buffer is full."""
        if self.used + channels > self.buf.shape[3]:
            return None
        v = self.buf[:, :, :, self.used:self.used + channels]
        self.used += channels
        return v


class BlockInputFn(torch.autograd.Function):
    """Places the block input into the first channels of the block buffer; gradients pass straight through."""

    @staticmethod
    def forward(ctx, x, buf):
        xh = x.permute(0, 2, 3, 1)
        C = xh.shape[3]
        buf[:, :, :, :C].copy_(xh)
        return _nchw_view(buf[:, :, :, :C])

    @staticmethod
    def backward(ctx, g):
        return g, None


def _adjacent_channel_slices(tensors) -> bool:
    """True when the tensors are consecutive channel slices of ONE NHWC buffer (same pixel pitch, back to back)."""
    base = None
    for t in tensors:
        if t.dim() != 4 or t.dtype != torch.bfloat16:
            return False
        v = t.permute(0, 2, 3, 1)
        key = (t.untyped_storage().data_ptr(), v.stride(0), v.stride(1), v.stride(2), v.shape[0], v.shape[1], v.shape[2])
        if v.stride(3) != 1 and v.shape[3] != 1:
            return False
        if base is None:
            base, nxt = key, t.storage_offset() + v.shape[3]
            continue
        if key != base or t.storage_offset() != nxt:
            return False
        nxt += v.shape[3]
    return base is not None


class ConcatViewFn(torch.autograd.Function):
    """Channel concatenation of adjacent slices of one buffer: the result is a view, no bytes move.  Backward hands each
    input its channel range of the incoming gradient (views again; autograd sums the contributions of the later layers)."""

    @staticmethod
    def forward(ctx, *tensors):
        t0 = tensors[0]
        v0 = t0.permute(0, 2, 3, 1)
        N, H, W, _ = v0.shape
        ctx.splits = [t.shape[1] for t in tensors]
        C = sum(ctx.splits)
        return torch.as_strided(t0.detach(), (N, C, H, W), (v0.stride(0), 1, v0.stride(1), v0.stride(2)), t0.storage_offset())

    @staticmethod
    def backward(ctx, g):
        out, off = [], 0
        for c in ctx.splits:
            out.append(g.narrow(1, off, c))
            off += c
        return tuple(out)


class _GradSink:
    """Mailbox between a BnActFn (producer of a residual-branch gradient) and the ConvFn that consumes the same
    block input: the gradient is added inside that conv's dgrad epilogue instead of by a separate add kernel."""

    __slots__ = ("ptr", "pending", "consumed")

    def __init__(self, ptr):
        self.ptr, self.pending, self.consumed = ptr, None, False

    def offer(self, grad) -> bool:
        """Hand a gradient of the shared input to the sink's conv; False if that conv's backward already ran."""
        if self.consumed or self.pending is not None:
            return False
        self.pending = grad
        return True

    def take(self):
        t, self.pending = self.pending, None
        self.consumed = True
        return t


def _thin_group_width(conv) -> int:
    """Channels per group if ``conv`` is a thin-group convolution (ResNeXt: 4..16 channels per group, as many outputs
    as inputs per group) that runs as a 64-channel block-diagonal grouped conv on the tcgen05 path, else 0."""
    g = conv.groups
    if g == 1:
        return 0
    cin_g, cout_g = conv.in_channels // g, conv.out_channels // g
    if cin_g == cout_g and cin_g < 32 and 64 % cin_g == 0 and conv.in_channels % 64 == 0:
        return cin_g
    return 0


def _conv_weight(eng, conv):
    """bf16 KRSC weight and group count the kernels see.  Thin groups: the [K,R,S,cg] weight is scattered into a
    [K,R,S,64] block-diagonal one (zeros off the diagonal blocks) and the conv runs with C/64 groups of 64 channels --
    16x (cg=4) more MACs than strictly needed, but on the tensor cores instead of a library fallback."""
    cg = _thin_group_width(conv)
    w = eng.w16_krsc(conv.weight)
    if not cg:
        return w, conv.groups
    Kc, R, S, _ = w.shape
    dense = eng.scratch(f"blockdiag_w_{id(conv)}", (Kc, R, S, 64), torch.bfloat16)
    eng.K.blockdiag_pack(w, dense)
    return dense, conv.in_channels // 64


def _halo_ok(conv, xh, out) -> bool:
    """64 -> 64 channel 3x3 / stride 1 / pad 1 layers run on the halo-reuse kernel (conv3x3_halo.cu): the activation tile
    is loaded once per 128 output positions instead of once per filter tap.  B200_CONV_HALO=0 switches it off (A/B)."""
    import os
    return (out is None and conv.groups == 1 and conv.in_channels == 64 and conv.out_channels == 64
            and conv.kernel_size == (3, 3) and conv.stride == (1, 1) and conv.padding == (1, 1) and conv.dilation == (1, 1)
            and xh.shape[2] <= 62 and os.environ.get("B200_CONV_HALO", "1") != "0")


class ConvFn(torch.autograd.Function):
    """Implicit-GEMM convolution (fprop / dgrad / wgrad on tcgen05), optional BN-statistics epilogue."""

    @staticmethod
    def forward(ctx, x, eng, conv, stats, anchor, hand_to=None, out=None, peer=None):
        K = eng.K
        xh = _nhwc(x)
        w, groups = _conv_weight(eng, conv)
        N, H, W, C = xh.shape
        Kc, R, S, _ = w.shape
        s, p, d = conv.stride[0], conv.padding[0], conv.dilation[0]
        P = (H + 2 * p - d * (R - 1) - 1) // s + 1
        Q = (W + 2 * p - d * (S - 1) - 1) // s + 1
        # `out`: a channel slice of a wider NHWC buffer the epilogue stores into directly (dense blocks)
        y = out if (out is not None and tuple(out.shape) == (N, P, Q, Kc) and groups == 1) else \
            torch.empty((N, P, Q, Kc), dtype=torch.bfloat16, device=x.device)
        # peer (SyncBN): the kernel's last CTA announces this layer's statistics exchange to the other ranks
        ctx.halo = _halo_ok(conv, xh, out)
        if ctx.halo:
            K.conv3x3_halo(xh, w, y, stats, False, peer)
        else:
            K.conv_fprop(xh, w, y, stats, None, s, p, d, groups, peer)
        ctx.eng, ctx.conv = eng, conv
        ctx.save_for_backward(xh)
        ctx.x_needs_grad = x.requires_grad
        # Another consumer of this very input (a BnActFn whose residual it is, or a sibling conv of the block) may
        # hand us its gradient contribution: it is added inside our dgrad instead of by a separate add kernel.
        fusable = s == 1 or (s == 2 and d == 1 and R <= 5 and S <= 5)
        ctx.sink = _GradSink(xh.data_ptr()) if (x.requires_grad and fusable) else None
        eng.last_sink[conv] = ctx.sink
        # ... and we may hand OUR input gradient to a sibling conv's sink (it must read the same tensor)
        ctx.hand_to = hand_to if (hand_to is not None and x.requires_grad and hand_to.ptr == xh.data_ptr()) else None
        return _nchw_view(y)

    @staticmethod
    def backward(ctx, dy):
        eng, conv = ctx.eng, ctx.conv
        K = eng.K
        (xh,) = ctx.saved_tensors
        dyh = _nhwc(dy)
        s, p, d = conv.stride[0], conv.padding[0], conv.dilation[0]
        cg = _thin_group_width(conv)
        if cg:
            # weights were packed in forward (unchanged since: the bucket update is gated on mark_ready below)
            Kc, _, R, S = conv.weight.shape
            w, groups = eng.scratch(f"blockdiag_w_{id(conv)}", (Kc, R, S, 64), torch.bfloat16), conv.in_channels // 64
        else:
            w, groups = eng.w16_krsc(conv.weight), conv.groups

        def wgrad():
            if cg:
                dwd = eng.scratch("blockdiag_dw", (Kc, R, S, 64), torch.float32)
                dwd.zero_()
                K.conv_wgrad(dyh, xh, dwd, s, p, d, groups)
                K.blockdiag_unpack_add(dwd, eng.grad_krsc(conv.weight))
            else:
                K.conv_wgrad(dyh, xh, eng.grad_krsc(conv.weight), s, p, d, groups)
            # only now: the bucket's fused update overwrites the bf16 weights the dgrad (enqueued earlier) still reads
            eng.mark_ready(conv.weight)

        # SyncBN on several GPUs: the weight gradient is launched BETWEEN the two passes of the preceding layer's BN
        # backward (BnActFn.backward), where it hides that layer's cross-rank exchange; otherwise right here
        defer = eng.defer_wgrad and ctx.x_needs_grad
        dx = None
        if ctx.x_needs_grad:
            addend = ctx.sink.take() if ctx.sink is not None else None
            if s == 1:
                dxh = torch.empty_like(xh)
                if ctx.halo and addend is None:
                    K.conv3x3_halo(dyh, w, dxh, None, True, None)
                else:
                    K.conv_dgrad(dyh, w, dxh, 1, p, d, addend, groups)
            else:
                dxh = _strided_dgrad(K, dyh, w, xh.shape, s, p, d, groups, addend)
            if ctx.hand_to is not None and ctx.hand_to.offer(dxh):
                dx = None          # the sibling conv's dgrad adds it; autograd treats None as zero
            else:
                dx = _nchw_view(dxh)
        if defer:
            eng.defer(wgrad)
        else:
            wgrad()            # after the dgrad: mark_ready may launch the bucket update that rewrites these weights
        return dx, None, None, None, None, None, None, None


def _strided_dgrad(K, dyh, w, x_shape, s, p, d, groups=1, addend=None):
    """Data gradient of a strided conv.  Stride 2 (every strided layer of the zoo): ``conv_dgrad_s2`` computes the four
    (row, column) parity classes of dx as compact stride-1 tcgen05 dgrads over the taps that reach them and interleaves
    them in one pass -- the real FLOPs, no zero insertion, no ATen scatter.  ``addend``: a gradient of the same input
    from a sibling consumer; it is summed in the interleave pass, or -- 1x1 stride-2 projection -- updated in place at
    the sampled pixels only.  Other strides keep the zero-insertion fallback (stride^2 x the FLOPs)."""
    N, H, W, C = x_shape
    Kc, R, S, _ = w.shape
    if s == 2 and d == 1 and R <= 5 and S <= 5:
        if R == 1 and S == 1 and p == 0 and addend is not None:
            Pc, Qc = dyh.shape[1], dyh.shape[2]
            compact = torch.empty((N, Pc, Qc, C), dtype=dyh.dtype, device=dyh.device)
            K.conv_dgrad(dyh, w, compact, 1, 0, 1, None, groups)
            K.strided_add_inplace(addend, compact, s)
            return addend
        dxh = torch.empty((N, H, W, C), dtype=dyh.dtype, device=dyh.device)
        K.conv_dgrad_s2(dyh, w, dxh, p, addend, groups)
        return dxh
    assert addend is None
    if R == 1 and S == 1 and p == 0:
        Pc, Qc = dyh.shape[1], dyh.shape[2]
        compact = torch.empty((N, Pc, Qc, C), dtype=dyh.dtype, device=dyh.device)
        K.conv_dgrad(dyh, w, compact, 1, 0, 1, None, groups)
        dxh = torch.zeros((N, H, W, C), dtype=dyh.dtype, device=dyh.device)
        dxh[:, ::s, ::s, :][:, :Pc, :Qc] = compact
        return dxh
    P1 = H + 2 * p - d * (R - 1)
    Q1 = W + 2 * p - d * (S - 1)
    up = torch.zeros((N, P1, Q1, Kc), dtype=dyh.dtype, device=dyh.device)
    up[:, ::s, ::s, :][:, : dyh.shape[1], : dyh.shape[2]] = dyh
    dxh = torch.empty((N, H, W, C), dtype=dyh.dtype, device=dyh.device)
    K.conv_dgrad(up, w, dxh, 1, p, d, None, groups)
    return dxh


class DwConvFn(torch.autograd.Function):
    """Depthwise kxk convolution (k <= 5, stride 1/2) on the CUDA-core kernels of ``csrc/dwconv.cu``."""

    @staticmethod
    def forward(ctx, x, eng, conv, stats, anchor):
        K = eng.K
        xh = _nhwc(x)
        N, H, W, C = xh.shape
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        y = torch.empty((N, P, Q, C), dtype=torch.bfloat16, device=x.device)
        K.dw_fprop(xh, eng.w16_view(conv.weight).view(C, k, k), y, stats, k, s, p)
        ctx.eng, ctx.conv, ctx.geom = eng, conv, (k, s, p)
        ctx.save_for_backward(xh)
        ctx.x_needs_grad = x.requires_grad
        return _nchw_view(y)

    @staticmethod
    def backward(ctx, dy):
        eng, conv = ctx.eng, ctx.conv
        K = eng.K
        (xh,) = ctx.saved_tensors
        k, s, p = ctx.geom
        C = xh.shape[-1]
        dyh = _nhwc(dy)
        K.dw_wgrad(dyh, xh, eng.grad_flat_view(conv.weight).view(C, k, k), k, s, p)
        dx = None
        if ctx.x_needs_grad:
            dxh = torch.empty_like(xh)
            K.dw_dgrad(dyh, eng.w16_view(conv.weight).view(C, k, k), dxh, k, s, p)
            dx = _nchw_view(dxh)
        eng.mark_ready(conv.weight)
        return dx, None, None, None, None


class StemConvFn(torch.autograd.Function):
    """Stem convolution on the raw NCHW batch (fp32, already normalised, as the reference's loader produces,
    utils.py:127-135 -- or uint8 pixels, normalised with ``eng.input_mean/std`` while they are staged).
    The input needs no gradient.

    * 7x7 / stride 2 / pad 3 over 3 channels (ResNet, DenseNet, BoTNet): **space-to-depth**.  One pass turns the image
      into S[N, P+3, Q+3, 16] (2x2 pixel blocks as channels, csrc/extras.cu), on which the stem is a 4x4 stride-1
      convolution; four horizontally adjacent 16-channel pixels of S are 128 contiguous bytes, so an ``as_strided`` view
      with a W-stride of 16 elements presents them as ONE 64-channel pixel and the generic im2col tcgen05 kernels run
      fprop and wgrad as a 4x1 convolution over that (overlapping) view -- the im2col TMA map takes the strides as
      they are.  108 MB of S instead of 1.03 GB of explicit patches (written once, read by fprop and again by wgrad).
      For 64 output channels the forward has its own kernel (csrc/stem_conv.cu) that gathers the A tile with cp.async, so
      the four-fold overlap hits in L1 instead of crossing the L2 -> SM fabric four times.
    * any other small-Cin stem (3x3/2 of RegNet / EfficientNet): explicit im2col to [pixels, Kpad] bf16, then the GEMM."""

    @staticmethod
    def _s2d_ok(eng, conv, x):
        return (eng.stem_s2d and conv.kernel_size == (7, 7) and conv.stride == (2, 2) and conv.padding == (3, 3)
                and conv.in_channels == 3 and conv.out_channels % 8 == 0)

    @staticmethod
    def forward(ctx, x_nchw_f32, eng, conv, stats, anchor):
        K = eng.K
        N, C, H, W = x_nchw_f32.shape
        Kc, _, R, S = conv.weight.shape
        s, p = conv.stride[0], conv.padding[0]
        P = (H + 2 * p - (R - 1) - 1) // s + 1
        Q = (W + 2 * p - (S - 1) - 1) // s + 1
        x = x_nchw_f32.contiguous()
        norm = (list(eng.input_mean), list(eng.input_std)) if x.dtype == torch.uint8 else ()
        ctx.eng, ctx.conv = eng, conv
        if StemConvFn._s2d_ok(eng, conv, x):
            try:
                xs = torch.empty((N, P + 3, Q + 3, 16), dtype=torch.bfloat16, device=x.device)
                K.stem_s2d(x, xs, *norm)
                xv = StemConvFn._virtual(xs, Q)
                w2 = eng.scratch("stem_w_s2d", (Kc, 4, 1, 64), torch.bfloat16)
                K.stem_s2d_pack_w(eng.w16_krsc(conv.weight), w2)
                y = torch.empty((N, P, Q, Kc), dtype=torch.bfloat16, device=x.device)
                if eng.stem_gather and Kc == 64 and Q <= 128:
                    K.stem_conv_fprop(xs, w2, y, stats)      # A tile gathered with cp.async: each S row read once per tile
                else:
                    K.conv_fprop(xv, w2, y, stats, None, 1, 0, 1)
                ctx.s2d, ctx.Q = True, Q
                ctx.save_for_backward(xs)
                return _nchw_view(y)
            except RuntimeError as exc:          # e.g. a driver that rejects the overlapping tensor map: say so, keep training
                from loguru import logger
                logger.warning(f"[b200] space-to-depth stem unavailable ({exc}); using the explicit im2col stem")
                eng.stem_s2d = False
        kdim = R * S * C
        kpad = (kdim + 15) // 16 * 16
        patches = torch.empty((N * P * Q, 1, 1, kpad), dtype=torch.bfloat16, device=x.device)
        K.stem_im2col(x, patches, R, S, s, p, P, Q, *norm)
        wpad = eng.scratch("stem_w", (Kc, 1, 1, kpad), torch.bfloat16)
        K.pad_rows(eng.w16_krsc(conv.weight), wpad, Kc, kdim, kpad)
        y = torch.empty((N * P * Q, 1, 1, Kc), dtype=torch.bfloat16, device=x.device)
        K.conv_fprop(patches, wpad, y, stats, None, 1, 0, 1)
        ctx.s2d, ctx.dims = False, (Kc, kdim, kpad)
        ctx.save_for_backward(patches)
        return _nchw_view(y.view(N, P, Q, Kc))

    @staticmethod
    def _virtual(xs, Q):
        """[N, Hs, Ws, 16] -> the overlapping [N, Hs, Q, 64] view: pixel q = S-pixels q..q+3 (Ws = Q + 3)."""
        N, Hs, Ws, _ = xs.shape
        return xs.as_strided((N, Hs, Q, 64), (Hs * Ws * 16, Ws * 16, 16, 1))

    @staticmethod
    def backward(ctx, dy):
        eng, conv = ctx.eng, ctx.conv
        K = eng.K
        Kc = conv.out_channels
        (saved,) = ctx.saved_tensors
        if ctx.s2d:
            dwp = eng.scratch("stem_dw_s2d", (Kc, 4, 1, 64), torch.float32)
            dwp.zero_()
            K.conv_wgrad(_nhwc(dy), StemConvFn._virtual(saved, ctx.Q), dwp, 1, 0, 1)
            K.stem_s2d_unpack_dw(dwp, eng.grad_flat_view(conv.weight))
        else:
            Kc, kdim, kpad = ctx.dims
            dyh = _nhwc(dy).reshape(-1, 1, 1, Kc)
            dwp = eng.scratch("stem_dw", (Kc, 1, 1, kpad), torch.float32)
            dwp.zero_()
            K.conv_wgrad(dyh, saved, dwp, 1, 0, 1)
            K.unpad_add(dwp, eng.grad_flat_view(conv.weight), Kc, kdim, kpad)
        eng.mark_ready(conv.weight)
        return None, None, None, None, None


class BnActFn(torch.autograd.Function):
    """act(BN(y) + residual): one pass forward (statistics come from the conv epilogue or ``bn_stats``), two
    passes backward; SyncBN statistics are exchanged through peer memory inside the same kernels."""

    @staticmethod
    def forward(ctx, y, residual, eng, bn, act, stats_slot, training, anchor, sink=None):
        K = eng.K
        y2, (N, H, W, C) = _rows_view(y)
        res2 = _nhwc(residual).view(-1, C) if residual is not None else None
        out = torch.empty((N, H, W, C), dtype=torch.bfloat16, device=y.device)
        save = torch.empty((2, C), dtype=torch.float32, device=y.device)
        peer = eng.peer_state if (training and eng.sync_bn) else None
        count = float(y2.shape[0] * (eng.world if peer is not None else 1))
        stats = stats_slot.tensor if stats_slot is not None else save  # eval: unused
        # residual + ReLU layers: keep a 1-bit/element mask so backward need not re-read the residual (2 B/element,
        # twice) just to rebuild relu'(z)
        mask = None
        if training and residual is not None and act == "relu" and (y.requires_grad or residual.requires_grad):
            mask = torch.empty((y2.shape[0], C // 8), dtype=torch.uint8, device=y.device)
        presignaled = bool(peer is not None and stats_slot is not None and stats_slot.presignaled)
        K.bn_apply(y2, res2, out.view(-1, C), stats, stats_slot.sym_offset if stats_slot is not None else 0,
                   eng.master_view(bn.weight) if bn.affine else None, eng.master_view(bn.bias) if bn.affine else None,
                   bn.running_mean, bn.running_var, save[0], save[1], count, bn.eps,
                   bn.momentum if bn.momentum is not None else 0.1, ACT[act], training, peer, mask, presignaled)
        if stats_slot is not None:
            stats_slot.presignaled = False
        if training and bn.track_running_stats:
            eng.note_bn_step(bn)
        ctx.frozen = not training
        if ctx.frozen and (y.requires_grad or (residual is not None and residual.requires_grad)):
            # BN in eval mode inside a training step (frozen statistics): backward needs the constants it used
            save[0].copy_(bn.running_mean)
            save[1].copy_(torch.rsqrt(bn.running_var + bn.eps))
        ctx.eng, ctx.bn, ctx.act, ctx.count = eng, bn, act, count
        ctx.has_res = residual is not None
        ctx.res_needs_grad = residual is not None and residual.requires_grad
        # only fuse when the sink's conv really consumes this very tensor
        ctx.sink = sink if (sink is not None and residual is not None and ctx.res_needs_grad
                            and _nhwc(residual).data_ptr() == sink.ptr) else None
        need_res = residual is not None and act is not None and mask is None
        ctx.save_for_backward(y2, res2 if need_res else None, save, mask)
        ctx.shape = (N, H, W, C)
        return _nchw_view(out)

    @staticmethod
    def backward(ctx, dout):
        eng, bn = ctx.eng, ctx.bn
        K = eng.K
        y2, res2, save, mask = ctx.saved_tensors
        N, H, W, C = ctx.shape
        d2 = _nhwc(dout).view(-1, C)
        if ctx.frozen:
            return _frozen_bn_backward(ctx, y2, res2, save, d2)
        dy = torch.empty((N, H, W, C), dtype=torch.bfloat16, device=dout.device)
        dres = torch.empty((N, H, W, C), dtype=torch.bfloat16, device=dout.device) if ctx.res_needs_grad else None
        slot = eng.bwd_slot(bn)
        peer = eng.peer_state if eng.sync_bn else None
        args = (y2, d2, res2, dy.view(-1, C), dres.view(-1, C) if dres is not None else None, slot.tensor,
                slot.sym_offset, eng.master_view(bn.weight) if bn.affine else None,
                eng.master_view(bn.bias) if bn.affine else None, save[0], save[1],
                eng.grad_flat_view(bn.weight) if bn.affine else None,
                eng.grad_flat_view(bn.bias) if bn.affine else None, ctx.count, ACT[ctx.act], peer, mask)
        if peer is not None and eng.defer_wgrad:
            # reduce pass (its last CTA opens the cross-rank exchange) -> the weight-gradient GEMM of the conv that
            # consumed this BN's output, deferred by ConvFn.backward -> apply pass (waits for the peers' sums, which have
            # been travelling in the meantime)
            K.bn_backward(*args, 1)
            eng.flush_deferred()
            K.bn_backward(*args, 2)
        else:
            K.bn_backward(*args)
        if bn.affine:
            eng.mark_ready(bn.weight)
            eng.mark_ready(bn.bias)
        if ctx.sink is not None and dres is not None:
            ctx.sink.pending = dres      # consumed by the dgrad of the conv that reads the same block input
            dres = None
        return _nchw_view(dy), (_nchw_view(dres) if dres is not None else None), None, None, None, None, None, None, None


class BnReluPoolFn(torch.autograd.Function):
    """Stem tail ``maxpool3x3/2(relu(BN(y)))`` in one pass over the conv output, backward in two passes straight from the
    pooled gradient (csrc/elementwise.cu: bn_relu_pool_*): the normalised full-resolution map and its gradient -- 411 MB
    each for ResNet-50 at batch 256 -- are never written.  Reference: models/resnet.py:194-197 (three modules)."""

    @staticmethod
    def forward(ctx, y, eng, bn, stats_slot, training, anchor):
        K = eng.K
        yh = _nhwc(y)
        N, H, W, C = yh.shape
        out = torch.empty((N, H // 2, W // 2, C), dtype=torch.bfloat16, device=y.device)
        arg = torch.empty((N, H // 2, W // 2, C), dtype=torch.uint8, device=y.device) if (training and y.requires_grad) else None
        save = torch.empty((2, C), dtype=torch.float32, device=y.device)
        peer = eng.peer_state if (training and eng.sync_bn) else None
        count = float(N * H * W * (eng.world if peer is not None else 1))
        presignaled = bool(peer is not None and stats_slot is not None and stats_slot.presignaled)
        K.bn_relu_pool_fwd(yh, out, arg, stats_slot.tensor if stats_slot is not None else save,
                           stats_slot.sym_offset if stats_slot is not None else 0,
                           eng.master_view(bn.weight) if bn.affine else None, eng.master_view(bn.bias) if bn.affine else None,
                           bn.running_mean, bn.running_var, save[0], save[1], count, bn.eps,
                           bn.momentum if bn.momentum is not None else 0.1, training, peer, presignaled)
        if stats_slot is not None:
            stats_slot.presignaled = False
        if training and bn.track_running_stats:
            eng.note_bn_step(bn)
        ctx.eng, ctx.bn, ctx.count = eng, bn, count
        ctx.save_for_backward(yh, arg, save)
        return _nchw_view(out)

    @staticmethod
    def backward(ctx, dout):
        eng, bn = ctx.eng, ctx.bn
        K = eng.K
        yh, arg, save = ctx.saved_tensors
        dy = torch.empty_like(yh)
        slot = eng.bwd_slot(bn)
        peer = eng.peer_state if eng.sync_bn else None
        args = (yh, _nhwc(dout), arg, dy, slot.tensor, slot.sym_offset, eng.master_view(bn.weight) if bn.affine else None,
                save[0], save[1], eng.grad_flat_view(bn.weight) if bn.affine else None,
                eng.grad_flat_view(bn.bias) if bn.affine else None, ctx.count, peer)
        if peer is not None and eng.defer_wgrad:
            K.bn_relu_pool_bwd(*args, 1)       # its last CTA opens the cross-rank exchange
            eng.flush_deferred()               # a deferred weight-gradient GEMM hides the round trip
            K.bn_relu_pool_bwd(*args, 2)
        else:
            K.bn_relu_pool_bwd(*args, 3)
        if bn.affine:
            eng.mark_ready(bn.weight)
            eng.mark_ready(bn.bias)
        return _nchw_view(dy), None, None, None, None, None


def _frozen_bn_backward(ctx, y2, res2, save, d2):
    """Backward of BN applied with its running statistics (module in eval mode during training, e.g. fine-tuning
    with frozen BN): the statistics are constants, so dy = dz * gamma * invstd without the batch-mean corrections of
    the training-mode kernel.  Rare path, plain ATen ops on the saved bf16 tensors."""
    eng, bn = ctx.eng, ctx.bn
    N, H, W, C = ctx.shape
    mean, invstd = save[0], save[1]
    g = eng.master_view(bn.weight).detach().float() if bn.affine else torch.ones(C, device=y2.device)
    b = eng.master_view(bn.bias).detach().float() if bn.affine else torch.zeros(C, device=y2.device)
    xhat = (y2.float() - mean) * invstd
    dz = d2.float()
    if ctx.act is not None:
        z = xhat * g + b
        if res2 is not None:
            z = z + res2.float()
        if ctx.act == "relu":
            dz = dz * (z > 0)
        elif ctx.act == "silu":
            sg = torch.sigmoid(z)
            dz = dz * (sg * (1 + z * (1 - sg)))
        else:
            raise NotImplementedError(ctx.act)
    if bn.affine:
        eng.grad_flat_view(bn.weight).add_((dz * xhat).sum(0))
        eng.grad_flat_view(bn.bias).add_(dz.sum(0))
        eng.mark_ready(bn.weight)
        eng.mark_ready(bn.bias)
    dy = (dz * (g * invstd)).to(torch.bfloat16).view(N, H, W, C)
    dres = dz.to(torch.bfloat16).view(N, H, W, C) if ctx.res_needs_grad else None
    if ctx.sink is not None and dres is not None:
        ctx.sink.pending = dres
        dres = None
    return _nchw_view(dy), (_nchw_view(dres) if dres is not None else None), None, None, None, None, None, None, None


class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng, fc, anchor):
        K = eng.K
        x2 = x.contiguous()
        B, Cin = x2.shape
        w = eng.w16_view(fc.weight)  # [out, in]
        out = torch.empty((B, 1, 1, w.shape[0]), dtype=torch.bfloat16, device=x.device)
        K.conv_fprop(x2.view(B, 1, 1, Cin), w.view(w.shape[0], 1, 1, Cin), out, None,
                     eng.master_view(fc.bias) if fc.bias is not None else None, 1, 0, 1)
        ctx.eng, ctx.fc = eng, fc
        ctx.save_for_backward(x2)
        return out.view(B, -1)

    @staticmethod
    def backward(ctx, dout):
        eng, fc = ctx.eng, ctx.fc
        K = eng.K
        (x2,) = ctx.saved_tensors
        B, Cin = x2.shape
        d2 = dout.contiguous()
        Kc = d2.shape[1]
        K.conv_wgrad(d2.view(B, 1, 1, Kc), x2.view(B, 1, 1, Cin), eng.grad_flat_view(fc.weight).view(Kc, 1, 1, Cin), 1, 0, 1)
        if fc.bias is not None:
            K.colsum_add(d2, eng.grad_flat_view(fc.bias))
        dx = torch.empty((B, 1, 1, Cin), dtype=torch.bfloat16, device=dout.device)
        K.conv_dgrad(d2.view(B, 1, 1, Kc), eng.w16_view(fc.weight).view(Kc, 1, 1, Cin), dx, 1, 0, 1, None)
        eng.mark_ready(fc.weight)
        if fc.bias is not None:
            eng.mark_ready(fc.bias)
        return dx.view(B, Cin), None, None, None


class MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng, k, s, p):
        xh = _nhwc(x)
        N, H, W, C = xh.shape
        P, Q = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        out = torch.empty((N, P, Q, C), dtype=xh.dtype, device=x.device)
        arg = torch.empty((N, P, Q, C), dtype=torch.uint8, device=x.device) if x.requires_grad else None
        eng.K.maxpool_fwd(xh, out, arg, k, s, p)
        ctx.eng, ctx.cfg, ctx.in_shape = eng, (k, s, p), xh.shape
        ctx.save_for_backward(arg)
        return _nchw_view(out)

    @staticmethod
    def backward(ctx, dout):
        (arg,) = ctx.saved_tensors
        dx = torch.empty(ctx.in_shape, dtype=torch.bfloat16, device=dout.device)
        ctx.eng.K.maxpool_bwd(_nhwc(dout), arg, dx, *ctx.cfg)
        return _nchw_view(dx), None, None, None, None


class AvgPool2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng):
        xh = _nhwc(x)
        N, H, W, C = xh.shape
        out = torch.empty((N, H // 2, W // 2, C), dtype=xh.dtype, device=x.device)
        eng.K.avgpool2_fwd(xh, out)
        ctx.eng, ctx.in_shape = eng, xh.shape
        return _nchw_view(out)

    @staticmethod
    def backward(ctx, dout):
        dx = torch.empty(ctx.in_shape, dtype=torch.bfloat16, device=dout.device)
        ctx.eng.K.avgpool2_bwd(_nhwc(dout), dx)
        return _nchw_view(dx), None


class GapFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng):
        xh = _nhwc(x)
        N, H, W, C = xh.shape
        out = torch.empty((N, C), dtype=xh.dtype, device=x.device)
        eng.K.gap_fwd(xh, out)
        ctx.eng, ctx.in_shape = eng, xh.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        dx = torch.empty(ctx.in_shape, dtype=torch.bfloat16, device=dout.device)
        ctx.eng.K.gap_bwd(dout.contiguous(), dx)
        return _nchw_view(dx), None


class ChannelScaleFn(torch.autograd.Function):
    """x * gate[n, c] (squeeze-excite gating) with a one-pass backward (dx and the per-sample gate gradient)."""

    @staticmethod
    def forward(ctx, x, gate, eng):
        xh = _nhwc(x)
        g = gate.contiguous()
        out = torch.empty_like(xh)
        eng.K.channel_scale_fwd(xh, g, out)
        ctx.eng = eng
        ctx.save_for_backward(xh, g)
        return _nchw_view(out)

    @staticmethod
    def backward(ctx, dout):
        xh, g = ctx.saved_tensors
        dx = torch.empty_like(xh)
        dg = torch.zeros(g.shape, dtype=torch.float32, device=g.device)
        ctx.eng.K.channel_scale_bwd(_nhwc(dout), xh, g, dx, dg)
        return _nchw_view(dx), dg.to(g.dtype), None


class SeFn(torch.autograd.Function):
    """Squeeze-excite block as one autograd node: global average pool -> (fc1 + bias -> act -> fc2 + bias -> sigmoid)
    in ONE kernel for any squeeze width (EfficientNet's 4/6/10/20/28, RegNetY's 308, ...) -> channel scaling.
    Backward: one pass for dx = dout * gate and the gate gradient, the MLP backward (data part per sample group,
    weight part with one owner thread per weight: no atomics), and the pooled-input gradient broadcast-added into dx
    in place -- no ATen op, no full-size add of two dx tensors (reference reaches timm's SqueezeExcite)."""

    @staticmethod
    def forward(ctx, x, eng, fc1, fc2, act, anchor):
        K = eng.K
        xh = _nhwc(x)
        N, H, W, C = xh.shape
        r = fc1.out_channels
        pooled = torch.empty((N, C), dtype=torch.bfloat16, device=x.device)
        K.gap_fwd(xh, pooled)
        pre1 = torch.empty((N, r), dtype=torch.float32, device=x.device)
        gate = torch.empty((N, C), dtype=torch.bfloat16, device=x.device)
        K.se_gate_fwd(pooled, eng.w16_view(fc1.weight).view(r, C), eng.master_view(fc1.bias) if fc1.bias is not None else None,
                      eng.w16_view(fc2.weight).view(C, r), eng.master_view(fc2.bias) if fc2.bias is not None else None,
                      pre1, gate, ACT[act])
        out = torch.empty_like(xh)
        K.channel_scale_fwd(xh, gate, out)
        ctx.eng, ctx.fcs, ctx.act = eng, (fc1, fc2), act
        ctx.save_for_backward(xh, pooled, gate, pre1)
        return _nchw_view(out)

    @staticmethod
    def backward(ctx, dout):
        eng = ctx.eng
        K = eng.K
        fc1, fc2 = ctx.fcs
        xh, pooled, gate, pre1 = ctx.saved_tensors
        N, H, W, C = xh.shape
        r = fc1.out_channels
        dx = torch.empty_like(xh)
        dgate = torch.zeros((N, C), dtype=torch.float32, device=xh.device)
        K.channel_scale_bwd(_nhwc(dout), xh, gate, dx, dgate)
        ds = torch.empty((N, C), dtype=torch.float32, device=xh.device)
        scratch = eng.scratch("se_bwd", (N * C + 2 * N * r,), torch.float32)
        K.se_gate_bwd(dgate, gate, pooled, pre1, eng.w16_view(fc1.weight).view(r, C), eng.w16_view(fc2.weight).view(C, r),
                      eng.grad_flat_view(fc1.weight), eng.grad_flat_view(fc1.bias) if fc1.bias is not None else None,
                      eng.grad_flat_view(fc2.weight), eng.grad_flat_view(fc2.bias) if fc2.bias is not None else None,
                      ds, scratch, ACT[ctx.act])
        K.channel_add_bcast(dx, ds, 1.0 / (H * W))
        for fc in (fc1, fc2):
            eng.mark_ready(fc.weight)
            if fc.bias is not None:
                eng.mark_ready(fc.bias)
        return _nchw_view(dx), None, None, None, None, None


class MhsaFn(torch.autograd.Function):
    """BoTNet MHSA core (14x14 tokens, 128-wide heads) on the fused tcgen05 kernels of ``csrc/attention.cu``:
    forward = QK^T + relative-position logits + softmax + PV in one kernel on the NHWC projections; backward = two
    kernels (dQ with the softmax backward; dK / dV) plus the relative-table gradients as one grouped wgrad GEMM over all
    samples and heads.  Softmax probabilities are kept in bf16 for the backward pass (83 MB per layer at batch 256)."""

    @staticmethod
    def forward(ctx, qk, v, eng, rel_h, rel_w, heads, scale, anchor):
        K = eng.K
        qkh, vh = _nhwc(qk), _nhwc(v)
        B = qkh.shape[0]
        out = torch.empty_like(vh)
        p_save = torch.empty((B * heads, 196, 208), dtype=torch.bfloat16, device=qk.device)
        K.attn_fwd(qkh, vh, eng.w16_view(rel_w), eng.w16_view(rel_h), out, p_save, heads, scale)
        ctx.eng, ctx.rel, ctx.heads, ctx.scale = eng, (rel_h, rel_w), heads, scale
        ctx.save_for_backward(qkh, vh, p_save)
        return _nchw_view(out)

    @staticmethod
    def backward(ctx, dout):
        eng, heads = ctx.eng, ctx.heads
        K = eng.K
        rel_h, rel_w = ctx.rel
        qkh, vh, p_save = ctx.saved_tensors
        B = qkh.shape[0]
        dev = qkh.device
        douth = _nhwc(dout)
        dqk, dv = torch.empty_like(qkh), torch.empty_like(vh)
        ds_save = torch.empty((B * heads, 196, 256), dtype=torch.bfloat16, device=dev)
        dsrel = torch.empty((B, qkh.shape[1], qkh.shape[2], heads * 64), dtype=torch.bfloat16, device=dev)
        K.attn_bwd(douth, qkh, vh, eng.w16_view(rel_w), eng.w16_view(rel_h), p_save, ds_save, dsrel, dqk, dv, heads, ctx.scale)
        # d(rel tables)[m, :] = sum over samples, heads, queries of dS_rel[., m] * q[., :]: a grouped (per head) 1x1 wgrad
        dw = eng.scratch("mhsa_rel_dw", (heads * 64, 1, 1, 128), torch.float32)
        dw.zero_()
        K.conv_wgrad(dsrel, qkh[..., : heads * 128], dw, 1, 0, 1, heads)
        K.rel_grad_reduce(dw, eng.grad_flat_view(rel_w), eng.grad_flat_view(rel_h), heads)
        eng.mark_ready(rel_w)
        eng.mark_ready(rel_h)
        return _nchw_view(dqk), _nchw_view(dv), None, None, None, None, None, None


class CeTopkFn(torch.autograd.Function):
    """softmax-CE + top-1/top-k counts + dlogits in one kernel (reference trainer.py:43,50 + utils.py:265-277)."""

    @staticmethod
    def forward(ctx, logits, target, eng, topk):
        lg = logits.contiguous()
        B = lg.shape[0]
        accum = torch.zeros(3, dtype=torch.float32, device=lg.device)
        dl = torch.empty_like(lg) if logits.requires_grad else None
        eng.K.ce_topk(lg, target.contiguous(), dl, accum, topk, 1.0 / B)
        ctx.save_for_backward(dl)
        ctx.mark_non_differentiable(accum)
        return accum[0] / B, accum

    @staticmethod
    def backward(ctx, dloss, _daccum):
        (dl,) = ctx.saved_tensors
        return dl * dloss.to(dl.dtype), None, None, None


class _FcView:
    """Presents a 1x1 ``nn.Conv2d`` (weight [out, in, 1, 1], optional bias) to ``LinearFn`` as a linear layer."""

    def __init__(self, conv):
        self.weight, self.bias = conv.weight, conv.bias
        self.in_features, self.out_features = conv.in_channels, conv.out_channels


class NativeOps:
    """Implements the ``ops.functional`` surface for one ``NativeEngine``."""

    def __init__(self, engine):
        self.eng = engine
        # calls that left the sm_100a kernels for an ATen / library op, by kind; stays empty for every config/*.yaml model
        self.fallbacks: dict = {}

    def _fell_back(self, kind: str):
        self.fallbacks[kind] = self.fallbacks.get(kind, 0) + 1

    # ---- helpers -------------------------------------------------------------------------------
    @staticmethod
    def _native_conv_ok(conv: nn.Conv2d, x) -> bool:
        kh, kw = conv.kernel_size
        g = conv.groups
        cin_g, cout_g = conv.in_channels // g, conv.out_channels // g
        # grouped convs run as `g` independent implicit GEMMs in one launch; thin groups (ResNeXt's 4-16 channels)
        # are regrouped into 64-channel block-diagonal groups (``_conv_weight``)
        thin = _thin_group_width(conv) > 0
        group_ok = g == 1 or (cin_g >= 32 and cout_g >= 32) or thin
        width_ok = thin or (cin_g % 8 == 0 and cout_g % 8 == 0)
        return (group_ok and width_ok and conv.bias is None and kh == kw and conv.stride[0] == conv.stride[1]
                and conv.padding[0] == conv.padding[1] and conv.dilation[0] == conv.dilation[1]
                and isinstance(conv.padding, tuple) and x.dtype == torch.bfloat16 and conv.padding_mode == "zeros")

    @staticmethod
    def _depthwise_ok(conv: nn.Conv2d, x) -> bool:
        kh, kw = conv.kernel_size
        return (conv.groups == conv.in_channels == conv.out_channels and conv.groups > 1 and conv.bias is None
                and kh == kw and kh <= 5 and conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1]
                and conv.dilation == (1, 1) and conv.in_channels % 8 == 0 and x.dtype == torch.bfloat16)

    @staticmethod
    def _is_stem(conv: nn.Conv2d, x) -> bool:
        return (conv.in_channels < 8 and conv.groups == 1 and conv.bias is None and x.dtype in (torch.float32, torch.uint8)
                and conv.out_channels % 8 == 0 and conv.dilation[0] == 1)

    def _as_act(self, x):
        """Bring an arbitrary input (fp32 NCHW batch) onto the activation convention."""
        if x.dtype == torch.bfloat16:
            return x
        if x.dtype == torch.uint8:   # raw pixels reaching a non-stem op: normalise with ATen (rare path)
            from ..utils.data import normalize_uint8
            x = normalize_uint8(x, self.eng.input_mean, self.eng.input_std)
        N, C, H, W = x.shape
        out = torch.empty((N, H, W, C), dtype=torch.bfloat16, device=x.device)
        self.eng.K.nchw_to_nhwc(x.contiguous(), out)
        return _nchw_view(out)

    def _torch_conv(self, x, conv):
        self._fell_back("conv")
        x = self._as_act(x)
        w = self.eng.w16_leaf(conv.weight)
        b = self.eng.w16_leaf(conv.bias) if conv.bias is not None else None
        return F.conv2d(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)

    # ---- functional surface ----------------------------------------------------------------------
    def conv_bn_act(self, x, conv, bn, act, residual, residual_sink=None, input_grad_to=None, out_buffer=None):
        eng = self.eng
        training = bn is not None and bn.training
        slot = eng.fwd_slot(bn) if (training and conv.out_channels % 8 == 0) else None
        stats = slot.tensor if slot is not None else None
        if self._is_stem(conv, x):
            y = StemConvFn.apply(x, eng, conv, stats, eng.anchor)
        elif self._native_conv_ok(conv, x):
            hand_to = eng.last_sink.get(input_grad_to) if (input_grad_to is not None and torch.is_grad_enabled()) else None
            dest = out_buffer.take(conv.out_channels) if (out_buffer is not None and bn is None and act is None
                                                          and residual is None) else None
            peer = eng.peer_state if (slot is not None and eng.sync_bn) else None
            y = ConvFn.apply(x, eng, conv, stats, eng.anchor, hand_to, dest, peer)
            if peer is not None:
                slot.presignaled = True     # consumed (and cleared) by the BnActFn right below
        elif self._depthwise_ok(conv, x):
            y = DwConvFn.apply(x, eng, conv, stats, eng.anchor)
        else:
            y = self._torch_conv(x, conv)
            if stats is not None:
                yh = _nhwc(y)
                eng.K.bn_stats(yh.view(-1, yh.shape[-1]), stats)
        if bn is None:
            if residual is not None:
                y = y + residual
            return _torch_act(y, act)
        if y.shape[1] % 8 != 0:
            self._fell_back("batch_norm")
            w = eng.w16_leaf(bn.weight) if bn.affine else None
            b = eng.w16_leaf(bn.bias) if bn.affine else None
            y = F.batch_norm(y, bn.running_mean, bn.running_var, w, b, bn.training, bn.momentum or 0.1, bn.eps)
            if residual is not None:
                y = y + residual
            return _torch_act(y, act)
        sink = eng.last_sink.get(residual_sink) if (residual_sink is not None and torch.is_grad_enabled()) else None
        return BnActFn.apply(y, residual, eng, bn, act, slot, training, eng.anchor, sink)

    def conv_bn_relu_maxpool(self, x, conv, bn, k, s, p):
        """Stem: conv -> BN -> ReLU -> max-pool.  3x3/2/1 pooling of an even-sized map in training (or no-grad) mode runs
        the fused tail (BnReluPoolFn); everything else is the composition of the separate ops."""
        eng = self.eng
        stem = self._is_stem(conv, x)
        fused = ((k, s, p) == (3, 2, 1) and isinstance(bn, nn.BatchNorm2d) and conv.out_channels % 8 == 0
                 and (stem or self._native_conv_ok(conv, x)) and (bn.training or not torch.is_grad_enabled()))
        if fused:
            kh, st, pd, dl = conv.kernel_size[0], conv.stride[0], conv.padding[0], conv.dilation[0]
            ho = (x.shape[2] + 2 * pd - dl * (kh - 1) - 1) // st + 1
            wo = (x.shape[3] + 2 * pd - dl * (conv.kernel_size[1] - 1) - 1) // st + 1
            fused = ho % 2 == 0 and wo % 2 == 0 and ho >= 2 and wo >= 2
        if not fused:
            return self.max_pool2d(self.conv_bn_act(x, conv, bn, "relu", None), k, s, p)
        training = bn.training
        slot = eng.fwd_slot(bn) if training else None
        stats = slot.tensor if slot is not None else None
        if stem:
            y = StemConvFn.apply(x, eng, conv, stats, eng.anchor)
        else:
            peer = eng.peer_state if (slot is not None and eng.sync_bn) else None
            y = ConvFn.apply(self._as_act(x), eng, conv, stats, eng.anchor, None, None, peer)
            if peer is not None:
                slot.presignaled = True
        return BnReluPoolFn.apply(y, eng, bn, slot, training, eng.anchor)

    def bn_act(self, x, bn, act):
        eng = self.eng
        x = self._as_act(x)
        if x.shape[1] % 8 != 0:  # kernels are 8-channel vectorised; odd widths take the ATen path
            self._fell_back("batch_norm")
            w = eng.w16_leaf(bn.weight) if bn.affine else None
            b = eng.w16_leaf(bn.bias) if bn.affine else None
            y = F.batch_norm(x, bn.running_mean, bn.running_var, w, b, bn.training, bn.momentum or 0.1, bn.eps)
            return _torch_act(y, act)
        training = bn.training
        slot = eng.fwd_slot(bn) if training else None
        if slot is not None:
            x2, _ = _rows_view(x)
            eng.K.bn_stats(x2, slot.tensor)
        return BnActFn.apply(x, None, eng, bn, act, slot, training, eng.anchor)

    def linear(self, x, fc):
        if x.dtype == torch.bfloat16 and fc.in_features % 8 == 0 and fc.out_features % 8 == 0:
            return LinearFn.apply(x, self.eng, fc, self.eng.anchor)
        self._fell_back("linear")
        w = self.eng.w16_leaf(fc.weight)
        b = self.eng.w16_leaf(fc.bias) if fc.bias is not None else None
        return F.linear(x.to(torch.bfloat16), w, b)

    def max_pool2d(self, x, k, s, p):
        if x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0:
            return MaxPoolFn.apply(x, self.eng, k, s, p)
        self._fell_back("max_pool2d")
        return F.max_pool2d(x, k, s, p)

    def avg_pool2d(self, x, k, s):
        if x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0 and k == 2 and s == 2 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
            return AvgPool2Fn.apply(x, self.eng)
        self._fell_back("avg_pool2d")
        return F.avg_pool2d(x, k, s)

    def global_avg_pool(self, x):
        if x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0:
            return GapFn.apply(x, self.eng)
        self._fell_back("global_avg_pool")
        return x.mean(dim=(2, 3))

    def _se_fc(self, v, conv):
        """A squeeze-excite 1x1 conv on a pooled [N, C] vector: tcgen05 GEMM when the widths allow, else ATen."""
        cin, cout = conv.in_channels, conv.out_channels
        if cin % 8 == 0 and cout % 8 == 0 and conv.groups == 1:
            return LinearFn.apply(v, self.eng, _FcView(conv), self.eng.anchor)
        return self._torch_conv(v[:, :, None, None], conv).flatten(1)

    def squeeze_excite(self, x, fc1, fc2, act):
        if x.dtype != torch.bfloat16 or x.shape[1] % 8 != 0:
            self._fell_back("squeeze_excite")
            s = x.mean(dim=(2, 3), keepdim=True)
            s = _torch_act(self._torch_conv(s, fc1), act)
            return x * torch.sigmoid(self._torch_conv(s, fc2))
        if (fc1.groups == 1 and fc2.groups == 1 and fc1.kernel_size == (1, 1) and fc2.kernel_size == (1, 1)
                and x.shape[1] + fc1.out_channels <= 12000):
            return SeFn.apply(x, self.eng, fc1, fc2, act, self.eng.anchor)
        s = self.global_avg_pool(x)                                   # [N, C]  (native pooling kernel)
        h = _torch_act(self._se_fc(s, fc1), act)                      # [N, r]
        gate = torch.sigmoid(self._se_fc(h, fc2))                     # [N, C]
        return ChannelScaleFn.apply(x, gate, self.eng)

    def dense_block_buffer(self, x, total_channels):
        """(block input as a view into a fresh [N,H,W,total_channels] buffer, handle the block's convs store into)."""
        x = self._as_act(x)
        N, C, H, W = x.shape
        if C % 8 != 0 or total_channels % 8 != 0:
            return x, None
        buf = torch.empty((N, H, W, total_channels), dtype=torch.bfloat16, device=x.device)
        return BlockInputFn.apply(x, buf), _DenseBuffer(buf, C)

    def concat_channels(self, tensors):
        tensors = [self._as_act(t) for t in tensors]
        if len(tensors) == 1:
            return tensors[0]
        if _adjacent_channel_slices(tensors):
            return ConcatViewFn.apply(*tensors)
        self._fell_back("concat")
        return torch.cat(tensors, dim=1).contiguous(memory_format=torch.channels_last)

    def relpos_mhsa(self, qk, v, rel_h, rel_w, heads, dim_qk, dim_v, scale):
        B, _, H, W = qk.shape
        if (qk.dtype == torch.bfloat16 and v.dtype == torch.bfloat16 and H * W == 196 and H == 14 and dim_qk == 128
                and dim_v == 128 and tuple(rel_h.shape) == (27, 128) and tuple(rel_w.shape) == (27, 128)):
            return MhsaFn.apply(qk, v, self.eng, rel_h, rel_w, heads, float(scale), self.eng.anchor)
        self._fell_back("mhsa")
        from . import functional as Fn
        from . import runtime
        with runtime.native_scope(None):  # other map sizes / head widths: composite torch ops on the bf16 tensors
            return Fn._relpos_mhsa_composite(qk, v, self.eng.w16_leaf(rel_h), self.eng.w16_leaf(rel_w), heads, dim_qk, dim_v, scale)

    def relpos_attention(self, q, k, v, rel_h, rel_w, height, width, scale):
        self._fell_back("relpos_attention")
        from . import functional as Fn
        from . import runtime
        with runtime.native_scope(None):  # composite torch ops on bf16; positional tables via leaf views
            return Fn.relpos_attention(q, k, v, self.eng.w16_leaf(rel_h), self.eng.w16_leaf(rel_w), height, width, scale)

    def cross_entropy_topk(self, logits, target, topk):
        if logits.dtype == torch.bfloat16:
            loss, accum = CeTopkFn.apply(logits, target, self.eng, topk)
            return loss, accum[1], accum[2]
        from . import functional as Fn
        from . import runtime
        with runtime.native_scope(None):
            return Fn.cross_entropy_topk(logits, target, topk)


def _torch_act(x, act):
    if act is None:
        return x
    if act == "relu":
        return F.relu(x)
    if act == "silu":
        return F.silu(x)
    raise ValueError(act)
