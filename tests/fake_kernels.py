"""Plain-PyTorch emulation of the sm_100a kernel module (``distribuuuu_b200/_ext/b200_kernels``), same call
signatures, same memory conventions (bf16 NHWC activations, KRSC weights, fp32 flat gradients that are ACCUMULATED
into, [2][C] statistic slots, 1-bit ReLU masks, uint8 max-pool tap indices).

It lets the CPU test suite run the native engine's *host* logic end to end -- autograd Functions, gradient mailboxes
between block branches, strided-dgrad decompositions, BN slot bookkeeping, flat-buffer views, the fused optimizer
step -- and compare a training step against the fp32 torch path.  It says nothing about the CUDA kernels themselves;
those are checked against fp32 PyTorch on the GPU box (tests/test_gpu_kernels.py).
"""
import torch
import torch.nn.functional as F

BF16 = torch.bfloat16
ACT_NONE, ACT_RELU, ACT_SILU = 0, 1, 2


def _raw(t):
    """Alias of ``t`` with its own autograd version counter: kernels write through raw pointers, so a write into
    one region of the engine's big allocation (statistics slots, bf16 weights, staging share it) must not look like
    a modification of every other view of it."""
    return t.data


def _nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2).float()


def _act(z, act):
    return torch.relu(z) if act == ACT_RELU else (F.silu(z) if act == ACT_SILU else z)


def _act_grad(z, act):
    if act == ACT_RELU:
        return (z > 0).float()
    if act == ACT_SILU:
        s = torch.sigmoid(z)
        return s * (1 + z * (1 - s))
    return torch.ones_like(z)


def _pack_mask(positive):           # [rows, C] bool -> [rows, C/8] uint8, bit i = channel cv*8+i
    rows, C = positive.shape
    w = (2 ** torch.arange(8, dtype=torch.int32)).view(1, 1, 8)
    return (positive.view(rows, C // 8, 8).int() * w).sum(-1).to(torch.uint8)


def _unpack_mask(mask, C):
    bits = (mask.int().unsqueeze(-1) >> torch.arange(8, dtype=torch.int32)) & 1
    return bits.view(mask.shape[0], C).float()


class FakeKernels:
    calls: dict

    def __init__(self):
        self.calls = {}

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    # ---------------------------------------------------------------- layout / casts
    def cast_bf16(self, src, dst):
        _raw(dst).copy_(src.to(BF16))

    def nchw_to_nhwc(self, x, out):
        _raw(out).copy_(x.permute(0, 2, 3, 1).to(BF16))

    def stem_im2col(self, x, patches, R, S, stride, pad, P, Q, mean=(), std=()):
        self._count("stem_im2col")
        if x.dtype == torch.uint8:
            m = torch.tensor(mean, dtype=torch.float32).view(1, -1, 1, 1)
            s = torch.tensor(std, dtype=torch.float32).view(1, -1, 1, 1)
            x = (x.float() / 255.0 - m) / s
        N, C = x.shape[:2]
        cols = F.unfold(x.float(), (R, S), padding=pad, stride=stride)            # [N, C*R*S, L], ordered (c, r, s)
        cols = cols.view(N, C, R * S, P * Q).permute(0, 3, 2, 1).reshape(N * P * Q, R * S * C)   # -> (r, s, c)
        flat = _raw(patches).view(N * P * Q, -1)
        flat.zero_()
        flat[:, : R * S * C] = cols.to(BF16)

    def stem_s2d(self, x, out, mean=(), std=()):
        self._count("stem_s2d")
        if x.dtype == torch.uint8:
            m = torch.tensor(mean, dtype=torch.float32).view(1, -1, 1, 1)
            sd = torch.tensor(std, dtype=torch.float32).view(1, -1, 1, 1)
            x = (x.float() / 255.0 - m) / sd
        N, C, H, W = x.shape
        Hs, Ws = out.shape[1], out.shape[2]
        xp = torch.zeros((N, C, 2 * Hs, 2 * Ws), dtype=torch.float32)
        xp[:, :, 3:3 + H, 3:3 + W] = x.float()
        o = _raw(out)
        o.zero_()
        for u in range(2):
            for v in range(2):
                blk = xp[:, :, u::2, v::2].permute(0, 2, 3, 1)                      # [N, Hs, Ws, 3]
                o[..., (u * 2 + v) * 3:(u * 2 + v) * 3 + 3] = blk.to(BF16)

    def stem_conv_fprop(self, xs, w2, y, stats, peer=None):
        self._count("stem_conv_fprop")
        N, Hs, Ws, _ = xs.shape
        Q = Ws - 3
        xv = xs.as_strided((N, Hs, Q, 64), (Hs * Ws * 16, Ws * 16, 16, 1))
        self.conv_fprop(xv, w2, y, stats, None, 1, 0, 1)

    def stem_s2d_pack_w(self, w, wp):
        K = w.numel() // 147
        w7 = torch.zeros((K, 8, 8, 3), dtype=torch.float32)
        w7[:, :7, :7] = w.reshape(K, 7, 7, 3).float()
        o = torch.zeros((K, 4, 4, 16), dtype=torch.float32)
        for u in range(2):
            for v in range(2):
                o[..., (u * 2 + v) * 3:(u * 2 + v) * 3 + 3] = w7[:, u::2, v::2]
        _raw(wp).copy_(o.reshape(K, 4, 1, 64).to(BF16))

    def stem_s2d_unpack_dw(self, dwp, dw):
        K = dw.numel() // 147
        o = dwp.reshape(K, 4, 4, 16)
        g = torch.zeros((K, 8, 8, 3), dtype=torch.float32)
        for u in range(2):
            for v in range(2):
                g[:, u::2, v::2] = o[..., (u * 2 + v) * 3:(u * 2 + v) * 3 + 3]
        _raw(dw).view(K, 7, 7, 3).add_(g[:, :7, :7])

    def pad_rows(self, src, dst, rows, cols, cols_pad):
        d = _raw(dst).view(rows, cols_pad)
        d.zero_()
        d[:, :cols] = src.reshape(rows, cols)

    def unpad_add(self, src, dst, rows, cols, cols_pad):
        _raw(dst.view(rows, cols)).add_(src.view(rows, cols_pad)[:, :cols])

    # ---------------------------------------------------------------- convolutions (implicit GEMM on the GPU)
    def conv_fprop(self, x, w, y, stats, bias, stride, pad, dil, groups=1, peer=None):
        self._count("conv_fprop")
        out = F.conv2d(_nchw(x), w.permute(0, 3, 1, 2).float(), bias.float() if bias is not None else None,
                       stride, pad, dil, groups)
        yb = out.permute(0, 2, 3, 1).to(BF16)
        _raw(y).copy_(yb.reshape(y.shape))
        if stats is not None:
            K = w.shape[0]
            f = yb.float().reshape(-1, K)
            _raw(stats)[:K] += f.sum(0)
            _raw(stats)[K:2 * K] += (f * f).sum(0)

    def conv3x3_halo(self, x, w, y, stats=None, dgrad=False, peer=None):
        self._count("conv3x3_halo" + ("+dgrad" if dgrad else ""))
        if dgrad:
            return self.conv_dgrad(x, w, y, 1, 1, 1)
        return self.conv_fprop(x, w, y, stats, None, 1, 1, 1)

    def conv_dgrad(self, dy, w, dx, stride, pad, dil, addend=None, groups=1):
        self._count("conv_dgrad" + ("+addend" if addend is not None else ""))
        assert stride == 1
        g = F.conv_transpose2d(_nchw(dy), w.permute(0, 3, 1, 2).float(), None, 1, pad, 0, groups, dil)
        g = g.permute(0, 2, 3, 1)
        if addend is not None:
            g = g + addend.float().reshape(g.shape)
        _raw(dx).copy_(g.to(BF16).reshape(dx.shape))

    def conv_dgrad_s2(self, dy, w, dx, pad, addend=None, groups=1):
        self._count("conv_dgrad_s2" + ("+addend" if addend is not None else ""))
        N, H, W, C = dx.shape
        P, Q = dy.shape[1], dy.shape[2]
        R, S = w.shape[1], w.shape[2]
        oph, opw = H - ((P - 1) * 2 - 2 * pad + R), W - ((Q - 1) * 2 - 2 * pad + S)
        g = F.conv_transpose2d(_nchw(dy), w.permute(0, 3, 1, 2).float(), None, 2, pad, (oph, opw), groups, 1)
        g = g.permute(0, 2, 3, 1)
        if addend is not None:
            g = g + addend.float().reshape(g.shape)
        _raw(dx).copy_(g.to(BF16).reshape(dx.shape))

    def attn_fwd(self, qk, v, relw, relh, out, p_save, heads, scale):
        self._count("attn_fwd")
        B = qk.shape[0]
        d = 128
        x = qk.float().reshape(B, 196, 2 * heads, d)
        q, k = x[:, :, :heads].permute(0, 2, 1, 3), x[:, :, heads:].permute(0, 2, 1, 3)      # [B,h,196,d]
        vv = v.float().reshape(B, 196, heads, d).permute(0, 2, 1, 3)
        idx = torch.arange(14)
        tw = relw.float()[(idx[None, :] - idx[:, None]) + 13]      # [xi, xj, d]
        th = relh.float()[(idx[None, :] - idx[:, None]) + 13]      # [yi, yj, d]
        q5 = q.reshape(B, heads, 14, 14, d)
        lw = torch.einsum("bnyxd,xjd->bnyxj", q5, tw)
        lh = torch.einsum("bnyxd,yid->bnyxi", q5, th)
        pos = (lh[..., :, None] + lw[..., None, :]).reshape(B, heads, 196, 196)
        logits = scale * (q @ k.transpose(-1, -2) + pos)
        pr = torch.softmax(logits, dim=-1).to(BF16)
        ps = _raw(p_save).view(B, heads, 196, 208)
        ps.zero_()
        ps[..., :196] = pr
        o = (pr.float() @ vv).permute(0, 2, 1, 3).reshape(out.shape)
        _raw(out).copy_(o.to(BF16))

    def attn_bwd(self, dout, qk, v, relw, relh, p_save, ds_save, dsrel, dqk, dv, heads, scale):
        self._count("attn_bwd")
        B = qk.shape[0]
        d = 128
        x = qk.float().reshape(B, 196, 2 * heads, d)
        q, k = x[:, :, :heads].permute(0, 2, 1, 3), x[:, :, heads:].permute(0, 2, 1, 3)
        vv = v.float().reshape(B, 196, heads, d).permute(0, 2, 1, 3)
        do = dout.float().reshape(B, 196, heads, d).permute(0, 2, 1, 3)
        pr = p_save.float().view(B, heads, 196, 208)[..., :196]
        dp = do @ vv.transpose(-1, -2)
        ds = (scale * pr * (dp - (pr * dp).sum(-1, keepdim=True))).to(BF16).float()
        dvv = pr.transpose(-1, -2) @ do
        ds5 = ds.reshape(B, heads, 14, 14, 14, 14)                 # [b,h,yi,xi,yj,xj]
        gw = ds5.sum(4)                                            # [b,h,yi,xi,xj]
        gh = ds5.sum(5)                                            # [b,h,yi,xi,yj]
        rel = torch.zeros(B, heads, 14, 14, 64)
        idx = torch.arange(14)
        for xi in range(14):
            rel[:, :, :, xi, idx - xi + 13] = gw[:, :, :, xi, :]
        for yi in range(14):
            rel[:, :, yi, :, 32 + idx - yi + 13] = gh[:, :, yi, :, :]
        rel = rel.to(BF16)
        _raw(dsrel).copy_(rel.permute(0, 2, 3, 1, 4).reshape(dsrel.shape))
        relf = rel.float().reshape(B, heads, 196, 64)
        dq = ds @ k + relf[..., :27] @ relw.float() + relf[..., 32:59] @ relh.float()
        dk = ds.transpose(-1, -2) @ q
        dx = torch.cat([dq.permute(0, 2, 1, 3), dk.permute(0, 2, 1, 3)], dim=2).reshape(dqk.shape)
        _raw(dqk).copy_(dx.to(BF16))
        _raw(dv).copy_(dvv.permute(0, 2, 1, 3).reshape(dv.shape).to(BF16))

    def rel_grad_reduce(self, dw, grad_w, grad_h, heads):
        self._count("rel_grad_reduce")
        t = dw.float().reshape(heads, 64, 128).sum(0)
        _raw(grad_w).view(27, 128).add_(t[:27])
        _raw(grad_h).view(27, 128).add_(t[32:59])

    def strided_add_inplace(self, dx, compact, stride):
        self._count("strided_add_inplace")
        P, Q = compact.shape[1], compact.shape[2]
        v = _raw(dx)[:, ::stride, ::stride, :][:, :P, :Q]
        v.copy_((v.float() + compact.float()).to(BF16))

    def colsum_add(self, d, out):
        self._count("colsum_add")
        _raw(out)[: d.shape[1]] += d.float().sum(0)

    def blockdiag_pack(self, thin, dense):
        self._count("blockdiag_pack")
        K, R, S, cg = thin.shape
        d = torch.zeros(K, R, S, 64, dtype=BF16)
        for k in range(K):
            g = (k % 64) // cg
            d[k, :, :, g * cg:(g + 1) * cg] = thin[k]
        _raw(dense).copy_(d.reshape(dense.shape))

    def blockdiag_unpack_add(self, dense, thin):
        self._count("blockdiag_unpack_add")
        K, R, S, cg = thin.shape
        d = dense.reshape(K, R, S, 64)
        t = _raw(thin)
        for k in range(K):
            g = (k % 64) // cg
            t[k] += d[k, :, :, g * cg:(g + 1) * cg]

    def se_gate_fwd(self, sp, w1, b1, w2, b2, pre1, gate, act):
        self._count("se_gate_fwd")
        N, C = sp.shape
        r = w1.numel() // C
        z1 = sp.float() @ w1.float().reshape(r, C).t() + (b1.float() if b1 is not None else 0)
        _raw(pre1).copy_(z1.reshape(pre1.shape))
        z2 = _act(z1, act) @ w2.float().reshape(C, r).t() + (b2.float() if b2 is not None else 0)
        _raw(gate).copy_(torch.sigmoid(z2).to(BF16))

    def se_gate_bwd(self, dgate, gate, sp, pre1, w1, w2, dw1, db1, dw2, db2, ds, scratch, act):
        self._count("se_gate_bwd")
        N, C = sp.shape
        r = w1.numel() // C
        g = gate.float()
        dz2 = dgate.float().reshape(N, C) * g * (1 - g)
        z1 = pre1.float().reshape(N, r)
        h = _act(z1, act)
        _raw(dw2).view(C, r).add_(dz2.t() @ h)
        if db2 is not None:
            _raw(db2).add_(dz2.sum(0))
        dz1 = (dz2 @ w2.float().reshape(C, r)) * _act_grad(z1, act)
        _raw(dw1).view(r, C).add_(dz1.t() @ sp.float())
        if db1 is not None:
            _raw(db1).add_(dz1.sum(0))
        _raw(ds).copy_((dz1 @ w1.float().reshape(r, C)).reshape(ds.shape))

    def channel_add_bcast(self, dx, ds, scale):
        self._count("channel_add_bcast")
        N, H, W, C = dx.shape
        _raw(dx).copy_((dx.float() + (ds.float().view(N, 1, 1, C) * scale)).to(BF16))

    def conv_wgrad(self, dy, x, dw, stride, pad, dil, groups=1):
        self._count("conv_wgrad")
        K, R, S, Cg = dw.shape
        gw = torch.nn.grad.conv2d_weight(_nchw(x), (K, Cg, R, S), _nchw(dy), stride, pad, dil, groups)
        _raw(dw).add_(gw.permute(0, 2, 3, 1))

    # ---------------------------------------------------------------- depthwise (CUDA-core direct conv on the GPU)
    def dw_fprop(self, x, w, y, stats, k, s, p):
        self._count("dw_fprop")
        C = x.shape[3]
        out = F.conv2d(_nchw(x), w.float().view(C, 1, k, k), None, s, p, 1, C)
        yb = out.permute(0, 2, 3, 1).to(BF16)
        _raw(y).copy_(yb)
        if stats is not None:
            f = yb.float().reshape(-1, C)
            _raw(stats)[:C] += f.sum(0)
            _raw(stats)[C:2 * C] += (f * f).sum(0)

    def dw_dgrad(self, dy, w, dx, k, s, p):
        self._count("dw_dgrad")
        C = dy.shape[3]
        H, W = dx.shape[1], dx.shape[2]
        P, Q = dy.shape[1], dy.shape[2]
        oph, opw = H - ((P - 1) * s - 2 * p + k), W - ((Q - 1) * s - 2 * p + k)
        g = F.conv_transpose2d(_nchw(dy), w.float().view(C, 1, k, k), None, s, p, (oph, opw), C)
        _raw(dx).copy_(g.permute(0, 2, 3, 1).to(BF16))

    def dw_wgrad(self, dy, x, dw, k, s, p):
        self._count("dw_wgrad")
        C = x.shape[3]
        gw = torch.nn.grad.conv2d_weight(_nchw(x), (C, 1, k, k), _nchw(dy), s, p, 1, C)
        _raw(dw).add_(gw.view(C, k, k))

    # ---------------------------------------------------------------- batch norm
    def bn_stats(self, y2, stats):
        C = y2.shape[1]
        f = y2.float()
        _raw(stats)[:C] += f.sum(0)
        _raw(stats)[C:2 * C] += (f * f).sum(0)

    def bn_apply(self, y2, res2, out2, stats, sym_offset, gamma, beta, rm, rv, save_mean, save_invstd, count, eps,
                 momentum, act, training, peer, relu_mask=None, presignaled=False):
        self._count("bn_apply")
        assert peer is None
        C = y2.shape[1]
        yf = y2.float()
        if training:
            mean = stats[:C] / count
            var = (stats[C:2 * C] / count - mean * mean).clamp_min(0)
            invstd = torch.rsqrt(var + eps)
            _raw(save_mean).copy_(mean)
            _raw(save_invstd).copy_(invstd)
            if rm is not None:
                _raw(rm).mul_(1 - momentum).add_(momentum * mean)
                _raw(rv).mul_(1 - momentum).add_(momentum * var * (count / max(count - 1.0, 1.0)))
        else:
            mean, invstd = rm, torch.rsqrt(rv + eps)
        g = gamma if gamma is not None else torch.ones(C)
        b = beta if beta is not None else torch.zeros(C)
        z = (yf - mean) * invstd * g + b
        if res2 is not None:
            z = z + res2.float()
        o = _act(z, act)
        _raw(out2).copy_(o.to(BF16))
        if relu_mask is not None:
            _raw(relu_mask).copy_(_pack_mask(o > 0))

    def bn_backward(self, y, dout, residual, dy, dresidual, sums, sym_offset, gamma, beta, save_mean, save_invstd,
                    dgamma, dbeta, count, act, peer, relu_mask=None, phase=3):
        self._count("bn_backward" + ("+mask" if relu_mask is not None else ""))
        assert peer is None and phase == 3
        C = y.shape[1]
        yf, d = y.float(), dout.float()
        g = gamma if gamma is not None else torch.ones(C)
        b = beta if beta is not None else torch.zeros(C)
        xhat = (yf - save_mean) * save_invstd
        if relu_mask is not None:
            dz = d * _unpack_mask(relu_mask, C)
        elif act != ACT_NONE:
            z = xhat * g + b
            if residual is not None:
                z = z + residual.float()
            dz = d * _act_grad(z, act)
        else:
            dz = d
        s0, s1 = dz.sum(0), (dz * xhat).sum(0)
        _raw(sums)[:C] += s0
        _raw(sums)[C:2 * C] += s1
        if dgamma is not None:
            _raw(dgamma).add_(s1)
            _raw(dbeta).add_(s0)
        _raw(dy).copy_(((dz - s0 / count - xhat * (s1 / count)) * g * save_invstd).to(BF16))
        if dresidual is not None:
            _raw(dresidual).copy_(dz.to(BF16))

    # ---------------------------------------------------------------- pooling
    def maxpool_fwd(self, x, out, arg, k, s, p):
        self._count("maxpool_fwd")
        N, H, W, C = x.shape
        o, idx = F.max_pool2d(_nchw(x), k, s, p, return_indices=True)             # idx = h * W + w
        _raw(out).copy_(o.permute(0, 2, 3, 1).to(BF16))
        if arg is not None:
            P, Q = o.shape[2], o.shape[3]
            h, w = idx // W, idx % W
            ph = torch.arange(P).view(1, 1, P, 1)
            q = torch.arange(Q).view(1, 1, 1, Q)
            tap = (h - (ph * s - p)) * k + (w - (q * s - p))
            _raw(arg).copy_(tap.permute(0, 2, 3, 1).to(torch.uint8))

    def maxpool_bwd(self, dout, arg, dx, k, s, p):
        self._count("maxpool_bwd")
        N, H, W, C = dx.shape
        P, Q = dout.shape[1], dout.shape[2]
        tap = arg.long()
        ph = torch.arange(P).view(1, P, 1, 1)
        q = torch.arange(Q).view(1, 1, Q, 1)
        h = ph * s - p + tap // k
        w = q * s - p + tap % k
        n = torch.arange(N).view(N, 1, 1, 1).expand_as(tap)
        c = torch.arange(C).view(1, 1, 1, C).expand_as(tap)
        acc = torch.zeros((N, H, W, C), dtype=torch.float32)
        acc.index_put_((n, h, w, c), dout.float(), accumulate=True)
        _raw(dx).copy_(acc.to(BF16))

    # ---------------------------------------------------------------- fused stem tail: BN + ReLU + max-pool 3x3/2/1
    def bn_relu_pool_fwd(self, y, out, arg, stats, sym_offset, gamma, beta, rm, rv, save_mean, save_invstd, count, eps,
                         momentum, training, peer=None, presignaled=False):
        self._count("bn_relu_pool_fwd")
        assert peer is None
        N, H, W, C = y.shape
        assert H % 2 == 0 and W % 2 == 0 and C % 8 == 0 and tuple(out.shape) == (N, H // 2, W // 2, C)
        yf = y.float()
        if training:
            mean = stats[:C] / count
            var = (stats[C:2 * C] / count - mean * mean).clamp_min(0)
            invstd = torch.rsqrt(var + eps)
            _raw(save_mean).copy_(mean)
            _raw(save_invstd).copy_(invstd)
            if rm is not None:
                _raw(rm).mul_(1 - momentum).add_(momentum * mean)
                _raw(rv).mul_(1 - momentum).add_(momentum * var * (count / max(count - 1.0, 1.0)))
        else:
            mean, invstd = rm, torch.rsqrt(rv + eps)
        g = gamma if gamma is not None else torch.ones(C)
        b = beta if beta is not None else torch.zeros(C)
        scale = g * invstd
        z = yf * scale + (b - mean * scale)
        o, idx = F.max_pool2d(z.permute(0, 3, 1, 2), 3, 2, 1, return_indices=True)          # idx = h * W + w
        _raw(out).copy_(o.clamp_min(0).permute(0, 2, 3, 1).to(BF16))
        if arg is not None:
            P, Q = o.shape[2], o.shape[3]
            h, w = idx // W, idx % W
            ph = torch.arange(P).view(1, 1, P, 1)
            q = torch.arange(Q).view(1, 1, 1, Q)
            tap = (h - (ph * 2 - 1)) * 3 + (w - (q * 2 - 1))
            tap = torch.where(o > 0, tap, torch.full_like(tap, 9))                         # 9: no positive tap
            _raw(arg).copy_(tap.permute(0, 2, 3, 1).to(torch.uint8))

    def bn_relu_pool_bwd(self, y, dout, arg, dy, sums, sym_offset, gamma, save_mean, save_invstd, dgamma, dbeta, count,
                         peer=None, phase=3):
        self._count("bn_relu_pool_bwd")
        assert peer is None and phase == 3
        N, H, W, C = y.shape
        P, Q = dout.shape[1], dout.shape[2]
        tap = arg.long()
        live = tap < 9
        tap = tap.clamp_max(8)
        ph = torch.arange(P).view(1, P, 1, 1)
        q = torch.arange(Q).view(1, 1, Q, 1)
        h = (ph * 2 - 1 + tap // 3).clamp(0, H - 1)
        w = (q * 2 - 1 + tap % 3).clamp(0, W - 1)
        n = torch.arange(N).view(N, 1, 1, 1).expand_as(tap)
        c = torch.arange(C).view(1, 1, 1, C).expand_as(tap)
        dz = torch.zeros((N, H, W, C), dtype=torch.float32)
        dz.index_put_((n, h, w, c), dout.float() * live, accumulate=True)
        g = gamma if gamma is not None else torch.ones(C)
        xhat = (y.float() - save_mean) * save_invstd
        s0, s1 = dz.sum((0, 1, 2)), (dz * xhat).sum((0, 1, 2))
        _raw(sums)[:C] += s0
        _raw(sums)[C:2 * C] += s1
        if dgamma is not None:
            _raw(dgamma).add_(s1)
            _raw(dbeta).add_(s0)
        _raw(dy).copy_(((dz - s0 / count - xhat * (s1 / count)) * g * save_invstd).to(BF16))

    def avgpool2_fwd(self, x, out):
        _raw(out).copy_(F.avg_pool2d(_nchw(x), 2).permute(0, 2, 3, 1).to(BF16))

    def avgpool2_bwd(self, dout, dx):
        g = _nchw(dout).repeat_interleave(2, 2).repeat_interleave(2, 3) * 0.25
        full = torch.zeros(dx.shape[0], dx.shape[3], dx.shape[1], dx.shape[2])
        full[:, :, : g.shape[2], : g.shape[3]] = g
        _raw(dx).copy_(full.permute(0, 2, 3, 1).to(BF16))

    def gap_fwd(self, x, out):
        _raw(out).copy_(x.float().mean(dim=(1, 2)).to(BF16))

    def gap_bwd(self, dout, dx):
        N, H, W, C = dx.shape
        _raw(dx).copy_((dout.float() / (H * W)).view(N, 1, 1, C).expand(N, H, W, C).to(BF16))

    def channel_scale_fwd(self, x, gate, out):
        _raw(out).copy_((x.float() * gate.float().view(gate.shape[0], 1, 1, -1)).to(BF16))

    def channel_scale_bwd(self, dout, x, gate, dx, dgate):
        d = dout.float()
        _raw(dx).copy_((d * gate.float().view(gate.shape[0], 1, 1, -1)).to(BF16))
        _raw(dgate).add_((d * x.float()).sum(dim=(1, 2)))

    # ---------------------------------------------------------------- loss / optimizer
    def ce_topk(self, logits, target, dlogits, accum, topk, grad_scale):
        self._count("ce_topk")
        lf = logits.float()
        logp = F.log_softmax(lf, dim=1)
        _raw(accum)[0] += -logp.gather(1, target.view(-1, 1)).sum()
        top = lf.topk(topk, dim=1).indices
        _raw(accum)[1] += (top[:, :1] == target.view(-1, 1)).any(1).float().sum()
        _raw(accum)[2] += (top == target.view(-1, 1)).any(1).float().sum()
        if dlogits is not None:
            g = logp.exp()
            g[torch.arange(lf.shape[0]), target] -= 1.0
            _raw(dlogits).copy_((g * grad_scale).to(BF16))

    def sgd_local(self, master, mom, grad, w16, off, n, lr, momentum, dampening, wd, nesterov, first, grad_scale, zero_grad):
        self._count("sgd_local")
        m, v, g = _raw(master)[off:off + n], _raw(mom)[off:off + n], _raw(grad)[off:off + n]
        d = g * grad_scale + wd * m
        if momentum != 0:
            if first:
                v.copy_(d)
            else:
                v.mul_(momentum).add_(d, alpha=1 - dampening)
            d = d + momentum * v if nesterov else v
        m.add_(d, alpha=-lr)
        if w16 is not None:
            _raw(w16[off:off + n]).copy_(m.to(BF16))
        if zero_grad:
            g.zero_()
