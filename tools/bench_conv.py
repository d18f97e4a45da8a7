#!/usr/bin/env python
"""Per-shape micro-benchmark of the tcgen05 implicit-GEMM kernels against cuDNN (bf16, channels_last) on the
23 unique ResNet-50 convolution shapes at per-GPU batch 256.  CUDA-event timing, L2 flushed
between iterations; writes gpurun_out/conv_shapes.json (copy the summary into profiles/)."""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# Cin, Cout, k, stride, Hin, count
SHAPES = [(64, 64, 1, 1, 56, 1), (64, 64, 3, 1, 56, 3), (64, 256, 1, 1, 56, 4), (256, 64, 1, 1, 56, 2),
          (256, 128, 1, 1, 56, 1), (128, 128, 3, 2, 56, 1), (128, 512, 1, 1, 28, 4), (256, 512, 1, 2, 56, 1),
          (512, 128, 1, 1, 28, 3), (128, 128, 3, 1, 28, 3), (512, 256, 1, 1, 28, 1), (256, 256, 3, 2, 28, 1),
          (256, 1024, 1, 1, 14, 6), (512, 1024, 1, 2, 28, 1), (1024, 256, 1, 1, 14, 5), (256, 256, 3, 1, 14, 5),
          (1024, 512, 1, 1, 14, 1), (512, 512, 3, 2, 14, 1), (512, 2048, 1, 1, 7, 3), (1024, 2048, 1, 2, 14, 1),
          (2048, 512, 1, 1, 7, 2), (512, 512, 3, 1, 7, 2)]


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
    ap.add_argument("--shapes", type=int, nargs="*", help="indices into SHAPES (default: all)")
    ap.add_argument("--no-cudnn", action="store_true", help="fprop of our kernel only")
    ap.add_argument("--no-stats", action="store_true", help="fprop without the BN-statistics epilogue")
    ap.add_argument("--ours-only", action="store_true", help="all three directions of our kernels, no cuDNN timing (A/B runs)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "conv_shapes.json"))
    a = ap.parse_args()
    from distribuuuu_b200.ops import build
    K = build.load()
    torch.backends.cudnn.benchmark = True
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    rows, tot = [], {"ours": 0.0, "cudnn": 0.0}
    B = a.batch
    shapes = [SHAPES[i] for i in a.shapes] if a.shapes else SHAPES
    for cin, cout, k, s, h, cnt in shapes:
        pad = k // 2
        p = (h + 2 * pad - k) // s + 1
        x = torch.randn(B, h, h, cin, device="cuda").to(torch.bfloat16)
        w = (torch.randn(cout, k, k, cin, device="cuda") * (cin * k * k) ** -0.5).to(torch.bfloat16)
        y = torch.empty(B, p, p, cout, device="cuda", dtype=torch.bfloat16)
        dy = torch.randn(B, p, p, cout, device="cuda").to(torch.bfloat16)
        dx = torch.empty_like(x)
        dw = torch.zeros(cout, k, k, cin, device="cuda")
        st = torch.zeros(2 * cout, device="cuda")
        flops = 2.0 * B * p * p * cout * cin * k * k
        xc, wc, dyc = x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), dy.permute(0, 3, 1, 2)
        r = {"shape": f"{cin}->{cout} k{k} s{s} {h}->{p}", "count": cnt, "gflop": flops / 1e9}
        r["fprop_ours_ms"] = timeit(lambda: K.conv_fprop(x, w, y, None if a.no_stats else st, None, s, pad, 1), a.iters, flush)
        if a.no_cudnn:
            rows.append(r)
            continue
        r["wgrad_ours_ms"] = timeit(lambda: K.conv_wgrad(dy, x, dw, s, pad, 1), a.iters, flush)
        if s == 1:
            r["dgrad_ours_ms"] = timeit(lambda: K.conv_dgrad(dy, w, dx, 1, pad, 1), a.iters, flush)
        else:   # stride 2: parity-class dgrad (compact stride-1 dgrads + one interleave pass)
            r["dgrad_ours_ms"] = timeit(lambda: K.conv_dgrad_s2(dy, w, dx, pad), a.iters, flush)
        if cin == 64 and cout == 64 and k == 3 and s == 1:   # halo-reuse kernel (what the engine runs for this layer)
            r["fprop_halo_ms"] = timeit(lambda: K.conv3x3_halo(x, w, y, st, False, None), a.iters, flush)
            r["dgrad_halo_ms"] = timeit(lambda: K.conv3x3_halo(dy, w, dx, None, True, None), a.iters, flush)
        if not a.ours_only:
            r["fprop_cudnn_ms"] = timeit(lambda: F.conv2d(xc, wc, None, s, pad), a.iters, flush)
            r["wgrad_cudnn_ms"] = timeit(lambda: torch.ops.aten.convolution_backward(
                dyc, xc, wc, None, [s, s], [pad, pad], [1, 1], False, [0, 0], 1, [False, True, False]), a.iters, flush)
            r["dgrad_cudnn_ms"] = timeit(lambda: torch.ops.aten.convolution_backward(
                dyc, xc, wc, None, [s, s], [pad, pad], [1, 1], False, [0, 0], 1, [True, False, False]), a.iters, flush)
        for kind in ("fprop", "dgrad", "wgrad"):
            r[f"{kind}_ours_tflops"] = flops / r[f"{kind}_ours_ms"] / 1e9
            tot["ours"] += cnt * r[f"{kind}_ours_ms"]
            tot[kind + "_ours"] = tot.get(kind + "_ours", 0.0) + cnt * r[f"{kind}_ours_ms"]
            if not a.ours_only:
                r[f"{kind}_cudnn_tflops"] = flops / r[f"{kind}_cudnn_ms"] / 1e9
                tot["cudnn"] += cnt * r[f"{kind}_cudnn_ms"]
                tot[kind + "_cudnn"] = tot.get(kind + "_cudnn", 0.0) + cnt * r[f"{kind}_cudnn_ms"]
        bytes_min = 2.0 * (x.numel() + y.numel())
        r["fprop_ours_gbs_min"] = bytes_min / r["fprop_ours_ms"] / 1e6
        rows.append(r)
        print(json.dumps(r), flush=True)
        del x, w, y, dy, dx, dw
    out = {"batch": B, "rows": rows, "total_ms_weighted": tot, "peaks": peaks}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print("TOTAL (count-weighted, ms):", tot)


if __name__ == "__main__":
    main()
