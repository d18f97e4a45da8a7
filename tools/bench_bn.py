#!/usr/bin/env python
"""Per-shape micro-benchmark of the BatchNorm kernels (forward apply, backward reduce, backward apply) on the
ResNet-50 BN shapes at per-GPU batch 256.  CUDA-event timing, L2 flushed between iterations; reports the
effective HBM bandwidth (minimal bytes / time) against the measured copy peak.  Writes gpurun_out/bn_shapes.json.

kinds: plain   = conv -> BN -> ReLU           (act' recomputed from the conv output)
       mask    = conv -> BN -> +res -> ReLU   (1-bit ReLU mask stored by forward, dResidual emitted)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# C, H(=W), kind, count in ResNet-50
SHAPES = [(64, 112, "plain", 1), (64, 56, "plain", 6), (256, 56, "mask", 3), (256, 56, "none", 1),
          (128, 56, "plain", 1), (128, 28, "plain", 7), (512, 28, "mask", 4), (512, 28, "none", 1),
          (256, 28, "plain", 1), (256, 14, "plain", 11), (1024, 14, "mask", 6), (1024, 14, "none", 1),
          (512, 14, "plain", 1), (512, 7, "plain", 5), (2048, 7, "mask", 3), (2048, 7, "none", 1)]


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--shapes", type=int, nargs="*")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bn_shapes.json"))
    a = ap.parse_args()
    from distribuuuu_b200.ops import build
    from distribuuuu_b200.ops.native import ACT
    K = build.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    hbm = 6572.0
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out, tot = [], {"fwd": 0.0, "reduce": 0.0, "apply": 0.0, "ideal": 0.0}
    idx = a.shapes if a.shapes else range(len(SHAPES))
    for i in idx:
        C, H, kind, cnt = SHAPES[i]
        rows = a.batch * H * H
        g = torch.Generator(device="cuda").manual_seed(i)
        y = torch.randn(rows, C, device="cuda", generator=g).bfloat16()
        dout = torch.randn(rows, C, device="cuda", generator=g).bfloat16()
        res = torch.randn(rows, C, device="cuda", generator=g).bfloat16() if kind == "mask" else None
        o = torch.empty_like(y)
        dy = torch.empty_like(y)
        dres = torch.empty_like(y) if kind == "mask" else None
        mask = torch.empty((rows, C // 8), dtype=torch.uint8, device="cuda") if kind == "mask" else None
        gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.zeros(C, device="cuda")
        rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        save = torch.empty(2, C, device="cuda")
        stats, sums = torch.zeros(2 * C, device="cuda"), torch.zeros(2 * C, device="cuda")
        dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        K.bn_stats(y, stats)
        act = ACT[None] if kind == "none" else ACT["relu"]

        def fwd():
            K.bn_apply(y, res, o, stats, 0, gamma, beta, rm, rv, save[0], save[1], float(rows), 1e-5, 0.1, act, True, None, mask)

        def bwd(phase):
            K.bn_backward(y, dout, None, dy, dres, sums, 0, gamma, beta, save[0], save[1], dg, db, float(rows), act, None, mask, phase)

        fwd()
        t_f = timeit(fwd, a.iters, flush)
        t_r = timeit(lambda: bwd(1), a.iters, flush)
        t_a = timeit(lambda: bwd(2), a.iters, flush)
        tb = rows * C * 2 / 1e9   # GB per tensor pass
        mb = rows * C / 8 / 1e9 if kind == "mask" else 0.0
        b_f = tb * (3 if kind == "mask" else 2) + mb
        b_r = tb * 2 + mb
        b_a = tb * (4 if kind == "mask" else 3) + mb
        row = {"C": C, "H": H, "kind": kind, "count": cnt, "fwd_ms": t_f, "reduce_ms": t_r, "apply_ms": t_a,
               "fwd_gbps": b_f / t_f * 1e3, "reduce_gbps": b_r / t_r * 1e3, "apply_gbps": b_a / t_a * 1e3}
        out.append(row)
        tot["fwd"] += cnt * t_f
        tot["reduce"] += cnt * t_r
        tot["apply"] += cnt * t_a
        tot["ideal"] += cnt * (b_f + b_r + b_a) / hbm * 1e3
        print(f"C={C:5d} H={H:4d} {kind:5s} x{cnt:2d} | fwd {t_f*1e3:7.1f} us {row['fwd_gbps']:6.0f} GB/s | reduce {t_r*1e3:7.1f} us "
              f"{row['reduce_gbps']:6.0f} GB/s | apply {t_a*1e3:7.1f} us {row['apply_gbps']:6.0f} GB/s", flush=True)
    print(f"TOTAL per step (weighted): fwd {tot['fwd']:.3f} ms, reduce {tot['reduce']:.3f} ms, apply {tot['apply']:.3f} ms; "
          f"HBM-roofline total {tot['ideal']:.3f} ms (copy peak {hbm:.0f} GB/s)")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"batch": a.batch, "hbm_copy_gbps": hbm, "shapes": out, "totals_ms": tot}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
