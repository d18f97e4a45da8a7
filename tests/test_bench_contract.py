"""Host-side pieces of the bench / profiling tooling that need no GPU: the bench CLI contract when no device is visible,
per-arch batch resolution (BASELINE configs), the clock-sample parser, and the launch-list step detection of the roofline
report (which has to follow the stem's kernel names)."""
import importlib.util
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_prints_one_json_line_and_exits_zero_without_a_gpu():
    for extra in ([], ["--impl", "reference"], ["--impl", "reference", "--arch", "regnety_160"]):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"] + extra,
                           capture_output=True, text=True, timeout=300, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [l for l in r.stdout.splitlines() if l.strip()]
        assert len(lines) == 1, r.stdout
        d = json.loads(lines[0])
        assert d["impl"] == ("reference" if extra else "ours") and "unavailable" in d


def test_batch_resolution_follows_the_baseline_configs():
    bench = _load("bench.py", "bench_mod")
    ns = types.SimpleNamespace
    assert bench._resolve_batch(ns(batch=0, arch="resnet50")) == 256          # BASELINE.json headline config
    assert bench._resolve_batch(ns(batch=0, arch="botnet50")) == 32           # config/botnet50.yaml
    assert bench._resolve_batch(ns(batch=0, arch="efficientnet_b0")) == 64
    assert bench._resolve_batch(ns(batch=48, arch="resnet50")) == 48
    assert "ResNet-50" in bench._metric_name("resnet50") and "images/sec" in bench._metric_name("regnety_160")


def test_clock_sampler_parses_throttle_reasons():
    bench = _load("bench.py", "bench_mod2")
    cs = bench.ClockSampler(0)
    assert cs.stop()["reasons"] == ["nvidia-smi unavailable"]                  # never started
    cs.proc = types.SimpleNamespace(terminate=lambda: None, wait=lambda timeout=None: 0, kill=lambda: None)
    cs.samples = ["1965, 1965, 512.3, Not Active, Not Active, Not Active, Not Active",
                  "1800, 1965, 690.1, Not Active, Not Active, Not Active, Active",
                  "1965, 1965, 600.0, Not Active, Not Active, Not Active, Not Active", "garbage"]
    out = cs.stop()
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"] and out["samples"] == 3
    assert out["power_w_max"] == 690.1


def test_roofline_launch_table_takes_the_last_complete_step(tmp_path):
    roof = _load("tools/roofline.py", "roofline_mod")
    rows = ['"ID","Kernel Name","Metric Name","Metric Unit","Metric Value"']
    names = ["void b200::stem_s2d_kernel<float>(const T1 *)", "void b200::conv_gemm_kernel<64, 0, 1>(CUtensorMap_st)",
             "b200::bn_relu_pool_fwd_strip_kernel(BnPoolParams)", "b200::sgd_local_kernel(float *)"]
    i = 0
    for step in range(3):                                   # two complete steps and a truncated third one
        for n in (names if step < 2 else names[:2]):
            rows.append(f'"{i}","{n}","gpu__time_duration.sum","ns","{1000 * (step + 1)}"')
            i += 1
    p = tmp_path / "launches.csv"
    p.write_text("==PROF== banner\n" + "\n".join(rows) + "\n")
    out = "\n".join(roof.launch_table(str(p)))
    assert "one training step = 4 kernel launches" in out and "0.01 ms" in out          # 4 x 2000 ns = step index 1
    assert "100.0%" in out.splitlines()[-1]                                                 # all b200:: kernels
