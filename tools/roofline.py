#!/usr/bin/env python
"""Turn the measured artefacts in profiles/ into a roofline report (profiles/ROOFLINE.md).

Inputs: a conv-shape table written by tools/bench_conv.py, a per-launch device-time list written by
`ncu --metrics gpu__time_duration.sum` around bench.py, and MEASURED_PEAKS.json (driver-measured denominators).
"""
import argparse
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p["hbm_gbs"], p["bf16_tflops"], p["bf16_tflops_sustained"], "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


def conv_table(path, hbm, tf_burst):
    d = json.load(open(path))
    B = d["batch"]
    lines = ["| shape (Cin->Cout k s H->P) | x | kind | ours ms | cuDNN ms | ours TFLOP/s | % of bf16 peak | min bytes GB/s | % of HBM peak | bound |",
             "|---|---|---|---|---|---|---|---|---|---|"]
    tot = collections.defaultdict(float)
    for r in d["rows"]:
        m = re.match(r"(\d+)->(\d+) k(\d) s(\d) (\d+)->(\d+)", r["shape"])
        cin, cout, k, s, h, p = map(int, m.groups())
        x_b, y_b, w_b = B * h * h * cin * 2, B * p * p * cout * 2, cout * cin * k * k * 2
        for kind in ("fprop", "dgrad", "wgrad"):
            if f"{kind}_ours_ms" not in r:
                continue
            ms, cms = r[f"{kind}_ours_ms"], r[f"{kind}_cudnn_ms"]
            nbytes = {"fprop": x_b + y_b + w_b, "dgrad": x_b + y_b + w_b, "wgrad": x_b + y_b + 2 * w_b}[kind]
            tfl = r["gflop"] / ms
            gbs = nbytes / ms / 1e6
            t_c, t_m = r["gflop"] / tf_burst, nbytes / hbm / 1e6
            bound = "compute" if t_c > t_m else "memory"
            frac = max(t_c, t_m) / ms
            lines.append(f"| {r['shape']} | {r['count']} | {kind} | {ms:.3f} | {cms:.3f} | {tfl:.0f} | {100 * tfl / tf_burst:.0f}% | "
                         f"{gbs:.0f} | {100 * gbs / hbm:.0f}% | {bound}: {100 * frac:.0f}% of roofline |")
            tot[kind + "_ours"] += r["count"] * ms
            tot[kind + "_cudnn"] += r["count"] * cms
    return lines, tot


def launch_table(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for row in csv.DictReader(lines):
        try:
            rows.append((row["Kernel Name"], float(row["Metric Value"].replace(",", ""))))
        except Exception:
            pass
    # one step = the stem's first kernel (space-to-depth / im2col) up to the optimizer kernel that follows it; take the LAST
    # COMPLETE one
    starts = [i for i, (n, v) in enumerate(rows) if "stem_im2col" in n or "stem_s2d_kernel" in n]
    ends = [i for i, (n, v) in enumerate(rows) if "sgd_local" in n or "allreduce_sgd" in n]
    step = rows
    for k in range(len(starts) - 1, -1, -1):
        limit = starts[k + 1] if k + 1 < len(starts) else len(rows)
        inside = [e for e in ends if starts[k] < e < limit]
        if inside:
            step = rows[starts[k]:inside[-1] + 1]
            break
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in step:
        short = re.sub(r"\(.*", "", n)
        short = re.sub(r"^void ", "", short)
        agg[short][0] += 1
        agg[short][1] += v * 1e-6
    total = sum(v[1] for v in agg.values())
    out = [f"one training step = {len(step)} kernel launches, {total:.2f} ms of device time (ncu serialises launches: compare shares)", "",
           "| kernel | launches | ms | share |", "|---|---|---|---|"]
    ours = 0.0
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
        out.append(f"| `{k[:90]}` | {c} | {t:.3f} | {100 * t / total:.1f}% |")
    for k, (c, t) in agg.items():
        if k.startswith("b200::"):
            ours += t
    out.append("")
    out.append(f"share of device time spent in this repo's kernels (`b200::*`): {100 * ours / total:.1f}%")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--conv", default=os.path.join(ROOT, "profiles", "conv_shapes_latest.json"))
    ap.add_argument("--launches", default=os.path.join(ROOT, "profiles", "launches_latest.csv"))
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "ROOFLINE.md"))
    a = ap.parse_args()
    hbm, tfb, tfs, src = load_peaks()
    md = [f"# Roofline report (denominators: {src} — HBM copy {hbm:.0f} GB/s, bf16 GEMM {tfb:.0f} TFLOP/s burst / {tfs:.0f} sustained)", ""]
    if os.path.exists(a.launches):
        md += ["## Where a ResNet-50 step goes (batch 256, 1 GPU)", ""] + launch_table(a.launches) + [""]
    if os.path.exists(a.conv):
        lines, tot = conv_table(a.conv, hbm, tfb)
        md += ["## tcgen05 implicit-GEMM kernels vs cuDNN, per ResNet-50 shape (batch 256, L2 flushed between iterations)", "",
               "roofline time = max(FLOPs / measured bf16 peak, minimal bytes / measured HBM bandwidth); "
               "`% of roofline` = roofline time / measured time.", ""] + lines + [""]
        md.append("count-weighted totals (ms per step): " + ", ".join(f"{k} {v:.2f}" for k, v in sorted(tot.items())))
    open(a.out, "w").write("\n".join(md) + "\n")
    print("wrote", a.out)


if __name__ == "__main__":
    main()
