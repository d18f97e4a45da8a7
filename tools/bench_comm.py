#!/usr/bin/env python
"""Micro-benchmark of the fused peer-memory all-reduce + SGD kernel on the ResNet-50 gradient buckets, against
NCCL all_reduce (+ the separate fused SGD kernel) on the same buffers.  Run under torchrun on N >= 2 GPUs:

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_comm.py

Reports per bucket: ms (device-timed, max over ranks), algorithm bandwidth S/t and bus bandwidth 2(N-1)/N * S/t
(NCCL convention, S = bf16 bytes of the bucket) and the fraction of the measured 770 GB/s per-direction NVLink peak.
Writes gpurun_out/comm_bench.json."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NVLINK_MEASURED_GBS = 770.0


def timed(fn, iters, dev):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def main():
    from distribuuuu_b200 import models, utils
    from distribuuuu_b200.parallel.native_engine import NativeEngine
    utils.setup_distributed()
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = utils.resolve_device()
    net = models.build_model("resnet50").to(dev)
    eng = NativeEngine(net, dev, sync_bn=False)
    eng.make_optimizer(lr=0.1, momentum=0.9, dampening=0.0, weight_decay=5e-5, nesterov=True)
    K = eng.K
    rows = []
    hyper = (0.1, 0.9, 0.0, 5e-5, True, False)
    for bi, b in enumerate(eng.buckets):
        nbytes16 = b.n * 2
        grid = 8 if b.one_shot else 120
        t_fused = timed(lambda: K.allreduce_sgd(eng.comm_state, eng.flat_master, eng.flat_mom, eng.flat_grad, b.off, b.n,
                                                *hyper, b.one_shot, grid), 20, dev)
        g = eng.flat_grad[b.off:b.off + b.n]

        def nccl_path():
            dist.all_reduce(g)
            K.sgd_local(eng.flat_master, eng.flat_mom, eng.flat_grad, eng.flat_w16, b.off, b.n, *hyper, 1.0 / world, True)

        t_nccl = timed(nccl_path, 20, dev)
        bus = 2.0 * (world - 1) / world * nbytes16 / (t_fused * 1e-3) / 1e9
        rows.append({"bucket": bi, "elements": b.n, "bf16_MiB": nbytes16 / 2 ** 20, "variant": "one-shot" if b.one_shot else "two-shot",
                     "fused_ms": t_fused, "nccl_fp32_allreduce_plus_sgd_ms": t_nccl, "algo_GBs": nbytes16 / (t_fused * 1e-3) / 1e9,
                     "bus_GBs": bus, "frac_of_770GBs": bus / NVLINK_MEASURED_GBS})
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)
    if rank == 0:
        out = {"world": world, "multicast": bool(getattr(eng, "has_multicast", False)), "buckets": rows,
               "total_fused_ms": sum(r["fused_ms"] for r in rows), "total_nccl_ms": sum(r["nccl_fp32_allreduce_plus_sgd_ms"] for r in rows)}
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"comm_bench_{world}gpu.json"), "w"), indent=1)
        print("TOTAL fused ms", out["total_fused_ms"], "nccl+sgd ms", out["total_nccl_ms"])
    utils.shutdown()


if __name__ == "__main__":
    main()
