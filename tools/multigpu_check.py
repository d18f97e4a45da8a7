#!/usr/bin/env python
"""Multi-GPU correctness of the peer-memory paths; run under torchrun on >= 2 GPUs of one box:

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multigpu_check.py

Checks (each prints PASS/FAIL on rank 0 and the script exits non-zero on any failure):
  1. allreduce_sgd (two-shot multicast, two-shot P2P, one-shot) == NCCL all-reduce + torch.optim.SGD
  2. peer-memory SyncBN forward/backward == BatchNorm over the concatenated global batch
  3. fp32 masters of BN gamma/beta/biases identical on every rank after several fused updates (no sync_masters)
  4. NativeEngine (peer comm, SyncBN) loss trajectory == TorchEngine (NCCL all_reduce, reference-semantics SyncBN)
  5. the multi-GPU step replayed from a CUDA graph == the same step launched eagerly
Results are also written to gpurun_out/multigpu_check.json.
"""
import copy
import json
import os
import sys
import traceback

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


_KEEP = []   # engines own symmetric-memory allocations; releasing them is left to process exit (every rank at once)


def make_engine(arch, dev, sync_bn, num_classes=16, cuda_graph=False):
    from distribuuuu_b200 import models
    from distribuuuu_b200.parallel.native_engine import NativeEngine
    torch.manual_seed(0)
    net = models.build_model(arch, num_classes=num_classes).to(dev)
    eng = NativeEngine(net, dev, sync_bn=sync_bn, cuda_graph=cuda_graph)
    _KEEP.append((net, eng))
    return net, eng


def stage(msg):
    """Per-rank progress line: if a world size hangs, the last line of every rank says where."""
    print(f"[rank {dist.get_rank()}] {msg}", file=sys.stderr, flush=True)


def check_allreduce_sgd(dev, rank, world):
    """Drive the fused kernel directly on a toy engine's flat buffers and compare with the library path."""
    net, eng = make_engine("resnet18", dev, sync_bn=False)
    assert eng.comm_mode == "peer", f"peer comm not active ({eng.comm_mode})"
    out = {"multicast": bool(eng.has_multicast)}
    K = eng.K
    variants = [("two_shot", False, True), ("two_shot_p2p", False, False), ("one_shot", True, False)]
    for name, one_shot, use_mc in variants:
        if use_mc and not eng.has_multicast:
            out[name] = "skipped (no multicast)"
            continue
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        n = eng.total
        master0 = torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
        eng.flat_master.copy_(master0)
        eng.flat_mom.zero_()
        ref_w = torch.nn.Parameter(master0.clone())
        opt = torch.optim.SGD([ref_w], lr=0.1, momentum=0.9, nesterov=True, weight_decay=5e-5)
        cs = eng.comm_state
        saved = (cs.mc_stage, cs.mc_w16)
        if not use_mc:
            cs.mc_stage, cs.mc_w16 = 0, 0
        for step in range(2):
            grad = torch.randn(n, device=dev, generator=g)
            eng.flat_grad.copy_(grad)
            gsum = grad.clone()
            dist.all_reduce(gsum)
            # the wire format is bf16 of (grad/world): mirror that rounding in the reference
            parts = [torch.empty_like(grad) for _ in range(world)]
            dist.all_gather(parts, grad)
            wire = sum((p / world).to(torch.bfloat16).float() for p in parts)
            if use_mc:  # multimem.ld_reduce accumulates in fp32 inside the switch but returns bf16x2
                wire = wire.to(torch.bfloat16).float()
            ref_w.grad = wire
            opt.step()
            torch.cuda.synchronize(dev)
            dist.barrier()
            K.allreduce_sgd(cs, eng.flat_master, eng.flat_mom, eng.flat_grad, 0, n, 0.1, 0.9, 0.0, 5e-5, True,
                            step == 0, one_shot, 32)
            torch.cuda.synchronize(dev)
            dist.barrier()
        cs.mc_stage, cs.mc_w16 = saved
        assert float(eng.flat_grad.abs().max()) == 0.0, "gradients not zeroed"
        if one_shot:
            lo, hi = 0, n
        else:
            per = (n // 8 + world - 1) // world
            lo, hi = min(per * rank, n // 8) * 8, min(per * rank + per, n // 8) * 8
        e_master = rel_err(eng.flat_master[lo:hi], ref_w.data[lo:hi])
        e_w16 = rel_err(eng.flat_w16[:n], ref_w.data)  # bf16 weights must be complete on EVERY rank
        out[name] = {"master_shard": e_master, "w16_all": e_w16}
        # multicast: the in-switch summation order may differ from ours by one bf16 ulp of the gradient
        assert e_master < (2e-3 if use_mc else 1e-5), f"{name}: master mismatch {e_master}"
        assert e_w16 < 1e-2, f"{name}: broadcast bf16 weights mismatch {e_w16}"
    return out


def check_syncbn(dev, rank, world):
    net, eng = make_engine("resnet18", dev, sync_bn=True)
    from distribuuuu_b200.ops.native import ACT
    K = eng.K
    C, rows = 64, 512
    torch.manual_seed(5)
    y_all = torch.randn(world * rows, C, device=dev).to(torch.bfloat16)
    d_all = torch.randn(world * rows, C, device=dev).to(torch.bfloat16)
    y, dout = y_all[rank * rows:(rank + 1) * rows].contiguous(), d_all[rank * rows:(rank + 1) * rows].contiguous()
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    bn = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d) and m.num_features == C][0]
    eng._begin_step()
    slot = eng.fwd_slot(bn)
    K.bn_stats(y, slot.tensor)
    out = torch.empty_like(y)
    save = torch.empty(2, C, device=dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    K.bn_apply(y, None, out, slot.tensor, slot.sym_offset, gamma, beta, rm, rv, save[0], save[1], float(world * rows),
               1e-5, 0.1, ACT["relu"], True, eng.peer_state)
    bslot = eng.bwd_slot(bn)
    dy = torch.empty_like(y)
    dgamma, dbeta = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    K.bn_backward(y, dout, None, dy, None, bslot.tensor, bslot.sym_offset, gamma, beta, save[0], save[1], dgamma, dbeta,
                  float(world * rows), ACT["relu"], eng.peer_state)
    torch.cuda.synchronize(dev)
    yr = y_all.float().clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    o = F.relu(F.batch_norm(yr, rm2, rv2, gr, br, True, 0.1, 1e-5))
    o.backward(d_all.float())
    sl = slice(rank * rows, (rank + 1) * rows)
    dg = dgamma.clone()
    dist.all_reduce(dg)
    errs = {"out": rel_err(out, o[sl]), "dy": rel_err(dy, yr.grad[sl]), "running_mean": rel_err(rm, rm2),
            "running_var": rel_err(rv, rv2), "dgamma_sum": rel_err(dg, gr.grad)}
    assert all(v < 3e-2 for v in errs.values()), errs
    return errs


def check_engine(dev, rank, world, arch="resnet18", steps=3, batch=16, size=64):
    from distribuuuu_b200.parallel import SyncBatchNorm
    from distribuuuu_b200.trainer import TorchEngine
    net_a, eng = make_engine(arch, dev, sync_bn=True)
    net_b = SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(net_a))
    # deepcopy keeps flat-view parameters; detach them into ordinary storage for the torch engine
    for p in net_b.parameters():
        p.data = p.data.clone().contiguous()
    opt = eng.make_optimizer(lr=0.01, momentum=0.9, dampening=0.0, weight_decay=5e-5, nesterov=True)
    ref = TorchEngine(net_b)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-5, nesterov=True)
    eng.train(), ref.train()
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    losses = []
    stage("engine: engines built")
    for i in range(steps):
        x = torch.randn(batch, 3, size, size, device=dev, generator=g)
        y = torch.randint(0, 16, (batch,), device=dev, generator=g)
        la, _, _ = eng.train_step(x, y, opt, 5)
        torch.cuda.synchronize(dev)      # the two engines never have kernels in flight at the same time
        stage(f"engine: native step {i} done")
        lb, _, _ = ref.train_step(x, y, ropt, 5)
        torch.cuda.synchronize(dev)
        stage(f"engine: reference step {i} done")
        losses.append((float(la), float(lb)))
    torch.cuda.synchronize(dev)
    # the loss of a step is the mean over ranks of different data; compare the all-reduced means
    lt = torch.tensor(losses, device=dev, dtype=torch.float64)
    dist.all_reduce(lt)
    losses = (lt / world).tolist()
    rel = max(abs(a - b) / max(abs(b), 1e-3) for a, b in losses)
    # every rank must hold identical bf16 weights after the fused updates
    mine = eng.flat_w16.float()
    other = mine.clone()
    dist.broadcast(other, src=0)
    same = float((mine - other).abs().max())
    # checkpoint path: gather sharded masters, compare with the bf16 weights
    sd = opt.state_dict()
    drift = rel_err(eng.flat_w16.float(), eng.flat_master)
    assert rel < 0.05, f"loss trajectories diverge {losses}"   # bf16 kernels vs the fp32 reference path; typically < 2 %
    assert same == 0.0, f"ranks disagree on weights by {same}"
    assert drift < 1e-2, f"master/bf16 mismatch after gather {drift}"
    assert len(sd["state"]) == len(eng.params)
    return {"losses": losses, "max_rel_loss_diff": rel, "rank_weight_diff": same, "master_vs_bf16": drift}


def check_fp32_masters(dev, rank, world, steps=4):
    """BN gamma/beta and biases are read in fp32 from the master buffer by the BN / bias kernels.
    After several fused updates -- WITHOUT sync_masters() -- every rank's fp32 copy of every 1-D parameter must be
    identical and must round to the broadcast bf16 weights; conv weights must also have moved (the update ran)."""
    net, eng = make_engine("resnet50", dev, sync_bn=True)
    opt = eng.make_optimizer(lr=0.05, momentum=0.9, dampening=0.0, weight_decay=5e-5, nesterov=True)
    eng.train()
    assert any(not b.one_shot for b in eng.buckets), "expected two-shot buckets for the conv weights"
    w0 = eng.flat_w16.float().clone()
    g = torch.Generator(device=dev).manual_seed(200 + rank)
    for _ in range(steps):
        x = torch.randn(8, 3, 64, 64, device=dev, generator=g)
        y = torch.randint(0, 16, (8,), device=dev, generator=g)
        eng.train_step(x, y, opt, 5)
    torch.cuda.synchronize(dev)
    lo, hi = eng.big_total, eng.trainable_total
    mine = eng.flat_master[lo:hi].clone()
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    diff_ranks = float((mine - ref).abs().max())
    vs_bf16 = rel_err(eng.flat_w16[lo:hi].float(), mine)
    moved_1d = float((mine - w0[lo:hi]).abs().max())
    moved_big = float((eng.flat_w16[:lo].float() - w0[:lo]).abs().max())
    assert diff_ranks == 0.0, f"fp32 masters of 1-D parameters differ across ranks by {diff_ranks}"
    assert vs_bf16 < 1e-2, f"fp32 masters of 1-D parameters do not match the bf16 weights ({vs_bf16})"
    assert moved_1d > 0 and moved_big > 0, "the fused update did not run"
    return {"rank_diff_1d_masters": diff_ranks, "masters_vs_bf16": vs_bf16, "moved_1d": moved_1d, "moved_big": moved_big,
            "syncbn_wait_ms_total": float(eng.syncbn_wait_ns.item()) / 1e6}


def check_graph_replay(dev, rank, world, arch="resnet18", steps=10, batch=16, size=64):
    """The multi-GPU training step captured in a CUDA graph (SyncBN exchanges + fused all-reduce on the side stream,
    device-side exchange counters) must follow the eagerly launched step: same data, same initial weights, two engines."""
    _, eager = make_engine(arch, dev, sync_bn=True)
    _, eager2 = make_engine(arch, dev, sync_bn=True)          # a second eager engine measures the run-to-run noise
    _, graphed = make_engine(arch, dev, sync_bn=True, cuda_graph=True)
    hp = dict(lr=0.005, momentum=0.9, dampening=0.0, weight_decay=5e-5, nesterov=True)
    oa, oc, ob = eager.make_optimizer(**hp), eager2.make_optimizer(**hp), graphed.make_optimizer(**hp)
    eager.train(), eager2.train(), graphed.train()
    noise = 0.0
    g = torch.Generator(device=dev).manual_seed(300 + rank)
    losses = []
    for i in range(steps):
        x = torch.randn(batch, 3, size, size, device=dev, generator=g)
        y = torch.randint(0, 16, (batch,), device=dev, generator=g)
        la, _, _ = eager.train_step(x, y, oa, 5)
        la = float(la)
        torch.cuda.synchronize(dev)
        lc, _, _ = eager2.train_step(x, y, oc, 5)
        noise = max(noise, abs(float(lc) - la) / max(abs(la), 1e-3))
        torch.cuda.synchronize(dev)
        lb, _, _ = graphed.train_step(x, y, ob, 5)
        lb = float(lb)
        torch.cuda.synchronize(dev)
        losses.append((la, lb))
        stage(f"graph: step {i} eager {la:.4f} graphed {lb:.4f} (replays so far {graphed.graph_replays})")
    assert graphed.graph_replays >= steps - 5, f"the step was not replayed from a graph ({graphed.graph_replays} replays)"
    rel = max(abs(a - b) / max(abs(a), 1e-3) for a, b in losses)
    mine = graphed.flat_w16.float()
    other = mine.clone()
    dist.broadcast(other, src=0)
    same = float((mine - other).abs().max())
    drift = rel_err(graphed.flat_w16.float(), eager.flat_w16.float())
    assert rel < max(0.02, 3 * noise), f"graph replay diverges from eager launches (eager-vs-eager noise {noise}) {losses}"
    assert same == 0.0, f"ranks disagree on weights after graph replays by {same}"
    return {"losses": losses, "max_rel_loss_diff": rel, "eager_vs_eager_noise": noise, "replays": graphed.graph_replays, "rank_weight_diff": same,
            "weights_vs_eager": drift}


def main():
    import faulthandler
    from distribuuuu_b200 import utils
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    faulthandler.dump_traceback_later(150, repeat=False, file=sys.stderr)   # a hang prints every rank's Python stack
    utils.setup_distributed()
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = utils.resolve_device()
    results, failed = {}, False
    for name, fn in [("allreduce_sgd", check_allreduce_sgd), ("syncbn", check_syncbn), ("fp32_masters", check_fp32_masters),
                     ("engine", check_engine), ("graph_replay", check_graph_replay)]:
        stage(f"{name}: start")
        try:
            results[name] = {"ok": True, "result": fn(dev, rank, world)}
        except Exception as exc:
            failed = True
            results[name] = {"ok": False, "err": f"{type(exc).__name__}: {exc}", "tb": traceback.format_exc()[-1500:]}
        flag = torch.tensor([0 if results[name]["ok"] else 1], device=dev)
        dist.all_reduce(flag)
        if rank == 0:
            print(("PASS " if flag.item() == 0 else "FAIL ") + name + " " + json.dumps(results[name])[:1200], flush=True)
        if flag.item() != 0:
            failed = True
            break  # a failed peer kernel may have poisoned the context
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump({"world": world, "results": results}, open(os.path.join(ROOT, "gpurun_out", "multigpu_check.json"), "w"), indent=1)
    try:
        utils.shutdown()
    except Exception:
        pass
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
