#!/usr/bin/env python
"""Headline benchmark: ResNet-50 training images/sec (BASELINE.json: resnet50.yaml, bf16, per-GPU batch 256,
SYNCBN=True, synthetic 3x224x224 data, random-init weights), device-timed, max over ranks.

    python bench.py                                  # 1 GPU, defaults finish in minutes
    torchrun --nproc-per-node N bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...             # the UNMODIFIED reference (baseline/_ref) on the same metric

Prints ONE JSON line on rank 0.  ``value`` is the whole-job throughput with inputs already on the device;
``e2e.value`` is the same metric through the public API (``engine.train_step``) with the host->device copy of
each step's pinned inputs and a device->host read of the loss inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch; 0 = 256 for resnet50 (BASELINE config), else the arch's "
                    "config/<arch>.yaml TRAIN.BATCH_SIZE")
    ap.add_argument("--arch", default="resnet50")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-variant", default="stock", choices=["stock", "amp"],
                    help="reference arm: as written (fp32 NCHW) or + autocast bf16 + channels_last")
    ap.add_argument("--no-syncbn", action="store_true")
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--e2e-input", default="both", choices=["both", "fp32", "uint8"],
                    help="dtype of the pinned host batches of the end-to-end run: uint8 = raw pixels normalised in the stem "
                         "kernel (B200.INPUT_UINT8, the framework's real-data path); fp32 = host-normalised images (what the "
                         "reference's loader yields); both = uint8 is reported as e2e, fp32 as e2e_fp32_input")
    ap.add_argument("--exposed", action="store_true", help="(kept for compatibility) the multi-GPU attribution runs -- step without gradient exchange, step without SyncBN -- are now always done when N > 1")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="CUDA-graph replay of the training step (B200.CUDA_GRAPH): auto = on")
    ap.add_argument("--attr-steps", type=int, default=10, help="steps of each multi-GPU attribution run (N > 1)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of this rank's GPU during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1])), power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------- helpers
def _dist_setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29512")
    os.environ.setdefault("RANK", "0"), os.environ.setdefault("WORLD_SIZE", "1"), os.environ.setdefault("LOCAL_RANK", "0")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    return rank, world, local, dev


def _timed(dev, fn, steps):
    """barrier + sync | K steps between CUDA events | sync + barrier; returns max-over-ranks seconds."""
    import torch
    import torch.distributed as dist
    dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    return float(ms.item()) / 1e3


BASELINE_PUBLISHED = None  # the reference publishes accuracy only -> vs_baseline is null

_PRETTY = {"resnet50": "ResNet-50", "regnety_160": "RegNetY-160", "regnetx_160": "RegNetX-160", "regnety_320": "RegNetY-320",
           "efficientnet_b0": "EfficientNet-B0", "botnet50": "BoTNet-50", "resnet18": "ResNet-18"}


def _metric_name(arch):
    return f"{_PRETTY.get(arch, arch)} training images/sec (whole job, device-timed, max over ranks)"


def _resolve_batch(args):
    if args.batch > 0:
        return args.batch
    if args.arch == "resnet50":
        return 256
    path = os.path.join(ROOT, "config", f"{args.arch}.yaml")
    try:
        import yaml
        return int(yaml.safe_load(open(path))["TRAIN"]["BATCH_SIZE"])
    except Exception:
        return 64


# --------------------------------------------------------------------------------------------- our arm
class _LaunchCounter:
    """Wraps the extension module and counts kernel launches of OUR kernels (per call -> kernels launched)."""
    PER_CALL = {"bn_backward": 2}

    def __init__(self, mod):
        self._m, self.count = mod, 0

    def __getattr__(self, name):
        attr = getattr(self._m, name)
        if not callable(attr) or isinstance(attr, type):
            return attr
        n = self.PER_CALL.get(name, 1)

        def wrapped(*a, **k):
            self.count += n
            return attr(*a, **k)
        return wrapped


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world, local, dev = _dist_setup()
    from distribuuuu_b200 import models
    from distribuuuu_b200.config import cfg
    from distribuuuu_b200.parallel.native_engine import NativeEngine
    torch.manual_seed(1)
    net = models.build_model(args.arch, num_classes=1000).to(dev)
    sync_bn = (not args.no_syncbn)
    use_graph = args.graph in ("on", "auto")
    eng = NativeEngine(net, dev, precision="bf16", comm=args.comm, bucket_cap_mb=cfg.B200.BUCKET_MB, sync_bn=sync_bn,
                       cuda_graph=use_graph)
    counter = _LaunchCounter(eng.K)
    eng.K = counter
    opt = eng.make_optimizer(lr=0.2, momentum=0.9, dampening=0.0, weight_decay=5e-5, nesterov=True)
    eng.train()
    B = args.batch = _resolve_batch(args)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    # a few distinct device-resident batches; each is 154 MB fp32 (> 126 MB L2), so inputs never sit in L2
    nbuf = 2
    xs = [torch.randn(B, 3, 224, 224, device=dev, generator=g) for _ in range(nbuf)]
    ys = [torch.randint(0, 1000, (B,), device=dev, generator=g) for _ in range(nbuf)]

    def step(i):
        eng.train_step(xs[i % nbuf], ys[i % nbuf], opt, 5)

    # with graph replay the capture happens on step 4 (three eager steps first): keep it inside the untimed warm-up
    for i in range(max(args.warmup, 8) if use_graph else args.warmup):
        step(i)
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local)
    sampler.start()
    counter.count = 0
    replays0 = eng.graph_replays
    if getattr(eng, "syncbn_wait_ns", None) is not None:
        eng.syncbn_wait_ns.zero_()
    sec = _timed(dev, step, args.steps)
    # kernels launched by Python calls + kernels launched by graph replays (counted once, at capture)
    launches = counter.count + (eng.graph_replays - replays0) * eng.graph_launches_per_step
    clocks = sampler.stop()
    value = world * B * args.steps / sec

    # Multi-GPU attribution (N > 1), outside the timed region above: the same step with (a) the gradient exchange
    # disabled and (b) SyncBN's cross-rank statistic exchange disabled, plus the device-side counter of the time the
    # designated CTAs spent inside the SyncBN exchanges during the timed region.
    exposed = None
    if world > 1:
        ksteps = max(3, min(args.attr_steps, args.steps))
        wait_ns = float(eng.syncbn_wait_ns.item()) if getattr(eng, "syncbn_wait_ns", None) is not None else None
        ms_full = sec * 1e3 / args.steps
        eng.debug_skip_comm = True
        for i in range(2):
            step(i)
        ms_nocomm = _timed(dev, step, ksteps) * 1e3 / ksteps
        eng.debug_skip_comm = False
        dist.broadcast(eng.flat_master, src=0)   # ranks diverged during the no-comm loop: restore a consistent state
        eng.refresh_compute_weights()
        ms_nosync = None
        if eng.sync_bn:
            eng.sync_bn = False
            for i in range(2):
                step(i)
            ms_nosync = _timed(dev, step, ksteps) * 1e3 / ksteps
            eng.sync_bn = True
            eng.sync_buffers()
        exposed = {"ms_per_step": ms_full, "ms_per_step_without_grad_exchange": ms_nocomm,
                   "exposed_allreduce_ms_per_step": ms_full - ms_nocomm,
                   "ms_per_step_without_syncbn_exchange": ms_nosync,
                   "syncbn_wait_ms_per_step": (ms_full - ms_nosync) if ms_nosync is not None else None,
                   "syncbn_exchange_device_ms_per_step": (wait_ns / 1e6 / args.steps) if wait_ns is not None else None,
                   "attr_steps": ksteps}

    e2e = e2e_alt = None
    if not args.skip_e2e:
        # the user-facing path: pinned host batches -> utils.PinnedPrefetcher (H2D of batch i+1 on a side stream while
        # step i computes, exactly what trainer.train_epoch does) -> engine.train_step -> loss read back every step
        from distribuuuu_b200 import utils as b200_utils
        hy = [torch.randint(0, 1000, (B,)).pin_memory() for _ in range(nbuf)]

        def measure_e2e(kind):
            if kind == "uint8":
                hx = [torch.randint(0, 256, (B, 3, 224, 224), dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
            else:
                hx = [torch.randn(B, 3, 224, 224).pin_memory() for _ in range(nbuf)]
            sink = []

            def run_e2e(n_steps):
                loader = b200_utils.PinnedPrefetcher([(hx[i % nbuf], hy[i % nbuf]) for i in range(n_steps)], dev)
                for x, y in loader:
                    loss, _, _ = eng.train_step(x, y, opt, 5)
                    sink.append(loss.item())                  # D2H read of the step's result

            run_e2e(max(3, args.warmup))
            dist.barrier()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run_e2e(args.steps)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
            sec_e2e = float(ms.item()) / 1e3
            note = ("raw uint8 pixels in pinned host memory, normalised inside the stem kernel (B200.INPUT_UINT8)" if kind == "uint8"
                    else "host-normalised fp32 images in pinned host memory (what the reference's loader yields)")
            return {"value": world * B * args.steps / sec_e2e, "unit": "images/sec",
                    "h2d_bytes_per_step": hx[0].numel() * hx[0].element_size() + B * 8, "d2h_bytes_per_step": 4,
                    "input": note, "ms_per_step": sec_e2e * 1e3 / args.steps, "last_loss": sink[-1] if sink else None}

        # primary e2e = host-normalised fp32 images, i.e. exactly what the reference arm is fed; the uint8 path (the
        # framework's own real-data path, 4x fewer PCIe bytes) is reported next to it
        if args.e2e_input == "both":
            e2e = measure_e2e("fp32")
            e2e_alt = measure_e2e("uint8")
        else:
            e2e = measure_e2e(args.e2e_input)
    if rank == 0:
        out = {"metric": _metric_name(args.arch), "impl": "ours",
               "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": sec * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": BASELINE_PUBLISHED, "dtype": "bf16", "data": "synthetic (random-init weights)",
               "config": {"model": args.arch, "global_batch": world * B, "per_gpu_batch": B, "image": "3x224x224",
                          "parallelism": f"dp{world}", "syncbn": bool(sync_bn and world > 1), "comm": eng.comm_mode,
                          "optimizer": "nesterov-sgd fused into the gradient all-reduce",
                          "cuda_graph": bool(eng.graph_replays > 0),
                          "l2": "inputs larger than L2 (154 MB fp32 per batch, alternating buffers)"},
               "clocks": clocks, "gpu_launches": launches, "gpu_launches_per_step": launches / max(args.steps, 1),
               "library_fallbacks": dict(eng.ops.fallbacks)}
        if e2e is not None:
            out["e2e"] = e2e
        if e2e_alt is not None:
            out["e2e_uint8_input"] = e2e_alt
        if exposed is not None:
            out["multi_gpu_attribution"] = exposed
            out["exposed_allreduce_ms"] = exposed["exposed_allreduce_ms_per_step"]
            out["syncbn_wait_ms"] = exposed["syncbn_wait_ms_per_step"]
        _emit(out)
    dist.barrier()
    dist.destroy_process_group()


# --------------------------------------------------------------------------------------------- reference arm
class _TimedLoader:
    """Feeds the reference's own ``train_epoch`` W+K pinned host batches and brackets the last K with CUDA
    events (start when batch W is handed out, end when the iterator is exhausted)."""

    def __init__(self, batches, warmup, dev):
        import torch
        self.batches, self.warmup, self.dev = batches, warmup, dev
        self.sampler = self
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.on_start = None

    def set_epoch(self, epoch):
        pass

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        import torch
        import torch.distributed as dist
        for i, b in enumerate(self.batches):
            if i == self.warmup:
                dist.barrier()
                torch.cuda.synchronize(self.dev)
                if self.on_start:
                    self.on_start()
                self.e0.record()
            yield b
        self.e1.record()
        torch.cuda.synchronize(self.dev)


def run_reference(args):
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "distribuuuu")):
        _emit({"impl": "reference", "unavailable": "baseline/_ref missing (run baseline/install_reference.sh)"})
        return
    sys.path.insert(0, os.path.join(ROOT, "baseline", "shims"))
    sys.path.insert(0, ref_dir)
    import torch
    import torch.distributed as dist
    import torch.nn as nn
    from torch.nn.parallel import DistributedDataParallel as DDP
    rank = int(os.environ.setdefault("RANK", "0"))
    world = int(os.environ.setdefault("WORLD_SIZE", "1"))
    local = int(os.environ.setdefault("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29513")
    try:
        from distribuuuu import models as ref_models, trainer as ref_trainer, utils as ref_utils
        from distribuuuu.config import cfg as rcfg
    except Exception as exc:
        _emit({"impl": "reference", "unavailable": f"import failed: {type(exc).__name__}: {exc}"})
        return
    args.batch = _resolve_batch(args)
    rcfg.MODEL.ARCH, rcfg.MODEL.SYNCBN, rcfg.MODEL.DUMMY_INPUT = args.arch, not args.no_syncbn, True
    rcfg.TRAIN.BATCH_SIZE, rcfg.TRAIN.PRINT_FREQ = args.batch, 10 ** 9
    rcfg.OUT_DIR = "/tmp/ref_bench_out"
    ref_utils.setup_distributed()                      # reference utils.py:19-51 (nccl, env://)
    dev = torch.device("cuda", local)
    ref_utils.setup_logger(rank, local)
    torch.backends.cudnn.benchmark = rcfg.CUDNN.BENCHMARK
    B, nb = args.batch, args.warmup + args.steps
    torch.manual_seed(100 + rank)
    pool = [(torch.randn(B, 3, 224, 224).pin_memory(), torch.randint(0, 1000, (B,)).pin_memory()) for _ in range(2)]

    def run_variant(variant):
        """The reference's own model / SyncBN conversion / DDP / optimizer / UNMODIFIED train_epoch.  ``amp`` is the
        precision-matched context arm: the very same loop under bf16 autocast with a channels_last model."""
        net = ref_models.build_model(arch=rcfg.MODEL.ARCH, pretrained=False, num_classes=rcfg.MODEL.NUM_CLASSES)
        net = nn.SyncBatchNorm.convert_sync_batchnorm(net) if rcfg.MODEL.SYNCBN else net   # trainer.py:131
        net = net.to(dev)
        if variant == "amp":
            net = net.to(memory_format=torch.channels_last)
        net = DDP(net, device_ids=[local], output_device=local)                              # trainer.py:134
        criterion = nn.CrossEntropyLoss().to(dev)
        optimizer = ref_utils.construct_optimizer(net)
        loader = _TimedLoader([pool[i % 2] for i in range(nb)], args.warmup, dev)
        sampler = ClockSampler(local)
        loader.on_start = sampler.start
        if variant == "amp":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ref_trainer.train_epoch(loader, net, criterion, optimizer, 0, 0, time.time())
        else:
            ref_trainer.train_epoch(loader, net, criterion, optimizer, 0, 0, time.time())   # UNMODIFIED hot loop
        clk = sampler.stop()
        ms = torch.tensor([loader.e0.elapsed_time(loader.e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sec_ = float(ms.item()) / 1e3
        del net, optimizer
        torch.cuda.empty_cache()
        return sec_, clk

    try:
        sec, clocks = run_variant(args.ref_variant)
    except Exception as exc:
        # e.g. regnety_160 / efficientnet_b0: the reference takes them from timm (trainer.py:124-128), which cannot be
        # installed offline -- the arm is unavailable for those configs, say so instead of substituting another model
        if rank == 0:
            _emit({"impl": "reference", "metric": _metric_name(args.arch), "n_gpus": world,
                   "unavailable": f"reference cannot build/run '{args.arch}' here: {type(exc).__name__}: {str(exc)[:200]}"})
        dist.barrier()
        dist.destroy_process_group()
        return
    value = world * B * args.steps / sec
    amp = None
    if args.ref_variant == "stock":
        # same box, same run: the reference loop at OUR compute precision (context for the headline ratio)
        try:
            sec_amp, clocks_amp = run_variant("amp")
            amp = {"value": world * B * args.steps / sec_amp, "unit": "images/sec", "ms_per_step": sec_amp * 1e3 / args.steps,
                   "dtype": "bf16 autocast + channels_last", "clocks": clocks_amp,
                   "note": "the reference's unmodified train_epoch under torch.autocast(bf16); not the reference as written"}
        except Exception as exc:  # never let the context arm break the stock number
            amp = {"unavailable": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        out = {"metric": _metric_name(args.arch), "impl": "reference",
               "variant": args.ref_variant, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": sec * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "fp32 (tf32 convs; reference as written)" if args.ref_variant == "stock" else "bf16 autocast",
               "data": "synthetic (random-init weights)",
               "config": {"model": args.arch, "global_batch": world * B, "per_gpu_batch": B, "parallelism": f"dp{world}",
                          "syncbn": bool(rcfg.MODEL.SYNCBN), "path": "distribuuuu.trainer.train_epoch (unmodified) + DDP/NCCL/cuDNN",
                          "l2": "inputs larger than L2 (154 MB per batch, pinned host -> device every step)"},
               "e2e": {"value": value, "unit": "images/sec", "h2d_bytes_per_step": B * 3 * 224 * 224 * 4 + B * 8,
                       "d2h_bytes_per_step": 12, "note": "the reference loop is end-to-end by construction"},
               "clocks": clocks, "gpu_launches": 0}
        if amp is not None:
            out["amp"] = amp
        _emit(out)
    dist.barrier()
    dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Library banners (e.g. 'NCCL version ...') go to fd 1; the contract is ONE JSON line on stdout, so fd 1 is
    pointed at stderr for the duration of the run and the JSON is written to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    args = parse_args()
    _quiet_stdout()
    import torch
    if not torch.cuda.is_available():
        _emit({"impl": args.impl, "unavailable": "no CUDA device visible (bench.py measures on B200)"})
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
