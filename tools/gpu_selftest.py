#!/usr/bin/env python
"""Run every kernel self-test in its own subprocess (timeout each) and write gpurun_out/selftest.json.
A trapping / hanging kernel therefore costs one check, not the whole GPU session.

    python tools/gpu_selftest.py            # all checks
    python tools/gpu_selftest.py --only fprop_3x3 bn_relu_res
    python tools/gpu_selftest.py --run NAME  # (internal) run one check in this process
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def registry():
    from distribuuuu_b200 import selftest as st
    checks = {name: (lambda n=name: st.check_conv_case(n)) for name in st.CONV_CASES}
    checks.update({
        "bn_relu_res": lambda: st.check_bn(act="relu", residual=True),
        "bn_relu_res_mask": lambda: st.check_bn(act="relu", residual=True, C=256, use_mask=True),
        "bn_none": lambda: st.check_bn(act=None, residual=False, C=256),
        "bn_silu_c24": lambda: st.check_bn(act="silu", residual=False, C=24),
        "pools": st.check_pools,
        "bn_relu_pool": st.check_bn_relu_pool,
        "bn_relu_pool_c96_odd": lambda: st.check_bn_relu_pool(N=3, H=10, W=14, C=96),
        "bn_relu_pool_prod": lambda: st.check_bn_relu_pool(N=64, H=112, W=112, C=64),
        "bn_relu_pool_wide_fallback": lambda: st.check_bn_relu_pool(N=2, H=4, W=600, C=64),
        "ce_topk": st.check_ce_topk,
        "sgd": st.check_sgd,
        "stem": st.check_stem,
        "stem_s2d": st.check_stem_s2d,
        "stem_s2d_224": lambda: st.check_stem_s2d(N=16, H=224, W=224),
        "stem_s2d_odd_batch_96": lambda: st.check_stem_s2d(N=3, H=96, W=160, Kc=96),
        "uint8_input": st.check_uint8_input,
        "channel_scale": st.check_channel_scale,
        "grouped_regnety": lambda: st.check_grouped_conv(C=224, K=224, G=2),
        "grouped_regnetx_s2": lambda: st.check_grouped_conv(C=512, K=512, G=4, stride=2, H=28, W=28),
        "grouped_232": lambda: st.check_grouped_conv(C=696, K=696, G=3, H=14, W=14),
        "dgrad_s2_3x3": lambda: st.check_dgrad_s2(),
        "dgrad_s2_3x3_odd_addend": lambda: st.check_dgrad_s2(N=3, H=7, W=7, C=64, K=192, with_addend=True),
        "dgrad_s2_1x1": lambda: st.check_dgrad_s2(N=2, H=28, W=28, C=256, K=512, R=1, pad=0),
        "dgrad_s2_1x1_addend": lambda: st.check_dgrad_s2(N=2, H=14, W=14, C=64, K=128, R=1, pad=0, with_addend=True),
        "dgrad_s2_grouped": lambda: st.check_dgrad_s2(N=2, H=28, W=28, C=224, K=224, G=2),
        "dgrad_s2_5x5": lambda: st.check_dgrad_s2(N=2, H=12, W=12, C=64, K=64, R=5, pad=2),
        "thin_groups_cg4": lambda: st.check_thin_groups(),
        "thin_groups_cg8_s2": lambda: st.check_thin_groups(C=256, G=32, H=28, W=28, stride=2),
        "thin_groups_cg16": lambda: st.check_thin_groups(C=512, G=32, H=7, W=7),
        "se_silu_r4": lambda: st.check_se(),
        "se_relu_r308": lambda: st.check_se(N=9, H=4, W=4, C=1232, r=308, act="relu"),
        "se_silu_r20": lambda: st.check_se(N=33, H=7, W=7, C=480, r=20),
        "colsum": st.check_colsum,
        "conv_halo_56": lambda: st.check_conv_halo(),
        "conv_halo_14_n3": lambda: st.check_conv_halo(N=3, H=14, W=14),
        "conv_halo_9x13": lambda: st.check_conv_halo(N=5, H=9, W=13),
        "conv_halo_prod": lambda: st.check_conv_halo(N=256, H=56, W=56),
        "mhsa": st.check_mhsa,
        "mhsa_b1": lambda: st.check_mhsa(B=1),
        "depthwise_5x5_s2": lambda: st.check_depthwise(k=5, stride=2),
        "depthwise_3x3_s1": lambda: st.check_depthwise(k=3, stride=1, C=32, H=16, W=16),
        "engine_resnet18": lambda: st.check_engine_vs_torch("resnet18", batch=16, size=64),
        "engine_resnet50": lambda: st.check_engine_vs_torch("resnet50", batch=8, size=64),
        "engine_resnext50": lambda: st.check_engine_vs_torch("resnext50_32x4d", batch=8, size=64, tol=0.15),
        "engine_densenet121": lambda: st.check_engine_vs_torch("densenet121", batch=8, size=64, tol=0.15),
        "engine_efficientnet_b0": lambda: st.check_engine_vs_torch("efficientnet_b0", batch=16, size=128, tol=0.15, lr=0.005),
        "engine_regnety_160": lambda: st.check_engine_vs_torch("regnety_160", batch=4, size=64, tol=0.15),
        "engine_regnetx_160": lambda: st.check_engine_vs_torch("regnetx_160", batch=4, size=64, tol=0.15),
        "engine_botnet50": lambda: st.check_engine_vs_torch("botnet50", batch=4, size=224, tol=0.15),
        "checkpoint_interop": st.check_checkpoint_interop,
        "grads_efficientnet_b0": lambda: st.check_engine_grads("efficientnet_b0"),
        "grads_resnet50": lambda: st.check_engine_grads("resnet50", batch=16, size=128),
        "grads_regnety_160": lambda: st.check_engine_grads("regnety_160", batch=8, size=128),
        "grads_densenet121": lambda: st.check_engine_grads("densenet121", batch=16, size=128),
        "grads_botnet50": lambda: st.check_engine_grads("botnet50", batch=4, size=224),
    })
    return checks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--run")
    ap.add_argument("--only", nargs="*")
    ap.add_argument("--timeout", type=int, default=180)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "selftest.json"))
    a = ap.parse_args()
    if a.run:
        res = registry()[a.run]()
        print("RESULT " + json.dumps(res))
        return
    names = a.only or list(registry())
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    results = {}
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--run", name], capture_output=True, text=True, timeout=a.timeout)
            ok = p.returncode == 0
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            results[name] = {"ok": ok, "sec": round(time.time() - t0, 1),
                             "result": json.loads(line[-1][7:]) if line else None,
                             "err": None if ok else (p.stderr[-1500:] + p.stdout[-500:])}
        except subprocess.TimeoutExpired:
            results[name] = {"ok": False, "sec": a.timeout, "err": "timeout"}
        print(f"{'PASS' if results[name]['ok'] else 'FAIL'} {name} {results[name].get('result')}", flush=True)
        if not results[name]["ok"]:
            print("   " + (results[name]["err"] or "").replace("\n", "\n   ")[-1200:], flush=True)
        with open(a.out, "w") as f:
            json.dump(results, f, indent=1)
    n_ok = sum(r["ok"] for r in results.values())
    print(f"{n_ok}/{len(results)} checks passed")


if __name__ == "__main__":
    main()
