"""Native data-parallel engine: flat parameter storage, fused peer-memory all-reduce + SGD,
peer-memory SyncBN, CUDA-stream overlap.

What replaces what (reference call sites -> here):

* ``DDP(net)`` ctor broadcast (trainer.py:134)      -> one broadcast of the flat fp32 master.
* C++ Reducer bucket copies + ``ncclAllReduce`` per bucket + ``optimizer.step()`` foreach kernels
  (trainer.py:46-47)                        -> ``csrc/comm.cu::allreduce_sgd_kernel`` per
  bucket on a side stream, started as soon as the bucket's last gradient kernel is enqueued: gradients are
  produced by the wgrad kernels directly into the flat fp32 buffer, reduced through NVLink peer memory
  (multimem.ld_reduce in the switch, or P2P loads), the Nesterov update runs on the owner's shard of the
  fp32 master weights (optimizer state is sharded ZeRO-1 style) and the new bf16 weights are multicast
  back to every rank.
* ``optimizer.zero_grad()`` (trainer.py:45; G14)                 -> folded into the same kernel.
* ``nn.SyncBatchNorm`` all_gather / all_reduce per layer (K5/K6) -> statistics exchanged by P2P loads inside
  the BN kernels (``csrc/elementwise.cu``), one flag exchange per layer per direction.
* per-iteration DDP buffer broadcast (K3)                        -> not per iteration: with SyncBN the running
  statistics are identical on every rank by construction; without it ``sync_buffers()`` broadcasts rank 0's
  before every validation / checkpoint, which is where the reference's broadcast becomes observable.
* parameters that are frozen (``requires_grad=False``) when the engine is built are kept out of the fused update,
  like ``torch.optim.SGD`` skips them; a *trainable* parameter that receives no gradient in some step still gets
  weight decay and momentum applied (the fused kernel runs over whole flat ranges) -- documented difference.

Checkpoints stay reference-compatible: ``module.state_dict()`` reads fp32 views of the flat master and
``FusedSGD.state_dict()`` emits ``torch.optim.SGD`` format (sharded state is gathered first).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List

import torch
import torch.distributed as dist
import torch.nn as nn
from loguru import logger

from ..ops import build, runtime
from ..ops.native import NativeOps

_ALIGN = 64           # parameter offsets are multiples of 64 elements (128 B of bf16: TMA-friendly)
_ONE_SHOT_BYTES = 512 * 1024


@dataclass
class _Slot:
    tensor: torch.Tensor
    sym_offset: int
    presignaled: bool = False   # the conv that filled the slot already announced the SyncBN exchange (its last CTA)


@dataclass
class _Bucket:
    off: int
    n: int
    params: list
    pending: int = 0
    one_shot: bool = False


class FusedSGD:
    """Optimizer facade over the engine's flat buffers (same surface the trainer and checkpoint code use on
    ``torch.optim.SGD``: ``param_groups``, ``state_dict``, ``load_state_dict``, ``zero_grad``, ``step``)."""

    def __init__(self, engine, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        self.engine = engine
        self.param_groups = [dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                  nesterov=nesterov, maximize=False, foreach=None, differentiable=False, fused=None,
                                  params=list(range(len(engine.params))))]
        self.steps = 0
        self.has_momentum_state = False

    def zero_grad(self, set_to_none: bool = True):  # gradients are zeroed by the fused update itself
        pass

    def step(self):  # the update is fused into the engine's train_step; kept for API compatibility
        pass

    def hyper(self):
        g = self.param_groups[0]
        return (float(g["lr"]), float(g["momentum"]), float(g["dampening"]), float(g["weight_decay"]),
                bool(g["nesterov"]), not self.has_momentum_state)

    def state_dict(self):
        eng = self.engine
        eng.sync_masters()
        state = {}
        if self.has_momentum_state and self.param_groups[0]["momentum"] != 0:
            for i, p in enumerate(eng.params):
                state[i] = {"momentum_buffer": eng.logical_view(eng.flat_mom, p).detach().clone().contiguous()}
        groups = [dict(g) for g in self.param_groups]
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        eng = self.engine
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(eng.params):
            raise ValueError("optimizer state does not match the model (parameter count differs)")
        for k in ("lr", "momentum", "dampening", "weight_decay", "nesterov"):
            if k in groups[0]:
                self.param_groups[0][k] = groups[0][k]
        eng.flat_mom.zero_()
        loaded = 0
        for i, p in enumerate(eng.params):
            st = sd["state"].get(i, sd["state"].get(str(i)))
            if st is not None and st.get("momentum_buffer") is not None:
                eng.logical_view(eng.flat_mom, p).copy_(st["momentum_buffer"])
                loaded += 1
        self.has_momentum_state = loaded > 0


class NativeEngine(nn.Module):
    def __init__(self, module: nn.Module, device: torch.device, precision: str = "bf16", comm: str = "peer",
                 bucket_cap_mb: float = 25, sync_bn: bool = False, cuda_graph: bool = False):
        super().__init__()
        # CUDA-graph replay of the training step (B200.CUDA_GRAPH); see _graphed_step
        self.cuda_graph = bool(cuda_graph)
        # 7x7/2 stems as a space-to-depth 4x1 convolution on the im2col tcgen05 kernels (ops/native.py: StemConvFn)
        self.stem_s2d = os.environ.get("B200_STEM_S2D", "1") != "0"
        self.stem_gather = os.environ.get("B200_STEM_GATHER", "0") == "1"    # csrc/stem_conv.cu (experiment, see its header)
        self._graphs = {}
        self._graph_pool = None
        self._eager_steps = 0
        self.graph_replays = 0
        self.graph_launches_per_step = 0
        if precision != "bf16":
            raise ValueError("the native engine computes in bf16 (fp32 master weights); use B200.ENGINE=torch for fp32")
        self.module = module
        self.device = device
        self.K = build.load()
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.sync_bn = bool(sync_bn) and self.world > 1
        self.ops = NativeOps(self)
        self.optimizer: FusedSGD | None = None
        self.comm_mode = comm if self.world > 1 else "local"
        self._scratch: Dict[str, torch.Tensor] = {}
        self._views: Dict[tuple, torch.Tensor] = {}
        self._slot_cache: Dict[tuple, "_Slot"] = {}
        self._leaves: Dict[nn.Parameter, torch.Tensor] = {}
        self._bn_stepped: List[torch.Tensor] = []
        self.last_sink: Dict[nn.Module, object] = {}   # conv module -> gradient mailbox of its latest forward
        self._slots_used = set()
        self._step_parity = 0
        self.comm_stream = torch.cuda.Stream(device)
        # autograd anchor: lets Functions whose tensor inputs carry no grad (first layer) still get a backward call
        self.anchor = torch.zeros((), device=device, requires_grad=True)
        # normalisation applied to uint8 image batches inside the stem kernel (B200.INPUT_UINT8); [0,1] units
        from ..utils.data import IMAGENET_MEAN, IMAGENET_STD
        self.input_mean, self.input_std = tuple(IMAGENET_MEAN), tuple(IMAGENET_STD)
        self._build_flat_storage()
        self._build_bn_slots()
        self._setup_comm()
        self._plan_buckets(int(bucket_cap_mb * 1024 * 1024))
        self._sync_initial_state()

    # ------------------------------------------------------------------------------ storage
    def _build_flat_storage(self):
        """Flat layout = [trainable weights with >= 2 dims | trainable 1-D parameters | frozen parameters], each
        region in registration order.

        Why the 1-D parameters (BN gamma/beta, biases) have their own region: the BN and bias kernels consume them
        in fp32 straight from the master buffer, and a two-shot bucket only keeps the *owner's* shard of the master
        up to date (every rank gets the new bf16 weights, not the fp32 ones).  Their region is therefore covered
        by one-shot buckets, where every rank reduces the whole bucket and updates its full fp32 replica.
        Frozen parameters (``requires_grad=False`` when the engine is built) sit past ``trainable_total`` and are
        never touched by the fused update -- ``torch.optim.SGD`` skips them too."""
        self.params = [p for p in self.module.parameters()]
        trainable = [p for p in self.params if p.requires_grad]
        self._region_big = [p for p in trainable if p.dim() >= 2]
        self._region_1d = [p for p in trainable if p.dim() < 2]
        frozen = [p for p in self.params if not p.requires_grad]
        self.index: Dict[nn.Parameter, tuple] = {}
        off = 0
        for region in (self._region_big, self._region_1d, frozen):
            for p in region:
                self.index[p] = (off, p.numel())
                off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            if region is self._region_big:
                self.big_total = off
            elif region is self._region_1d:
                self.trainable_total = off
        self.total = max(off, _ALIGN)
        dev = self.device
        self.flat_master = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.flat_mom = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(self.total, dtype=torch.float32, device=dev)
        for p in self.params:  # rebind every parameter to a view of the flat master (conv weights physically KRSC)
            view = self.logical_view(self.flat_master, p)
            view.copy_(p.data)
            p.data = view

    def logical_view(self, flat: torch.Tensor, p: nn.Parameter) -> torch.Tensor:
        """View of ``flat`` with the parameter's logical shape; 4-D weights are stored [O,H,W,I]."""
        off, n = self.index[p]
        if p.dim() == 4:
            O, I, H, W = p.shape
            return flat[off:off + n].view(O, H, W, I).permute(0, 3, 1, 2)
        return flat[off:off + n].view(p.shape)

    # The flat buffers are allocated once, so views into them are created once per parameter and reused: a
    # training step asks for ~600 of them, and slicing + reshaping in Python was a measurable part of the
    # per-step host time.
    def _cached(self, kind: str, p, make):
        views = self.__dict__.get("_views")
        if views is None:
            views = self.__dict__["_views"] = {}
        key = (kind, id(p))
        v = views.get(key)
        if v is None:
            v = views[key] = make()
        return v

    def master_view(self, p):
        return p.data

    def grad_flat_view(self, p):
        off, n = self.index[p]
        return self._cached("g", p, lambda: self.flat_grad[off:off + n])

    def grad_krsc(self, p):
        O, I, H, W = p.shape
        return self._cached("gk", p, lambda: self.grad_flat_view(p).view(O, H, W, I))

    def w16_view(self, p):
        off, n = self.index[p]
        return self._cached("w", p, lambda: self.flat_w16[off:off + n].view(p.shape))

    def w16_krsc(self, p):
        off, n = self.index[p]
        O, I, H, W = p.shape
        return self._cached("wk", p, lambda: self.flat_w16[off:off + n].view(O, H, W, I))

    def w16_leaf(self, p):
        """bf16 leaf (logical shape) for torch fallback ops; its gradient is folded into the flat buffer."""
        leaf = self._leaves.get(p)
        if leaf is None:
            leaf = self.logical_view(self.flat_w16, p).detach().requires_grad_(True)
            gview = self.logical_view(self.flat_grad, p)

            def hook(t, gview=gview, p=p):
                gview.add_(t.grad)
                t.grad = None
                self.mark_ready(p)

            leaf.register_post_accumulate_grad_hook(hook)
            self._leaves[p] = leaf
        return leaf

    def scratch(self, name, shape, dtype):
        t = self._scratch.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.zeros(shape, dtype=dtype, device=self.device)
            self._scratch[name] = t
        return t

    # ------------------------------------------------------------------------------ BN statistics slots
    def _build_bn_slots(self):
        self.bn_offsets: Dict[nn.Module, int] = {}
        off = 0
        for m in self.module.modules():
            if isinstance(m, nn.modules.batchnorm._BatchNorm):
                self.bn_offsets[m] = off
                off += 2 * m.num_features          # forward: sum, sumsq
                off += 2 * m.num_features          # backward: sum(dz), sum(dz*xhat)
        self.stats_len = (off + 63) // 64 * 64

    def _slot(self, bn, which: int) -> _Slot:
        ck = (id(bn), which, self._step_parity)
        slot = self._slot_cache.get(ck)
        if slot is None:
            n = 2 * bn.num_features
            base = self._step_parity * self.stats_len + self.bn_offsets[bn] + which * n
            slot = self._slot_cache[ck] = _Slot(self.stats_buf[base:base + n], self.stats_sym_base + base)
        # all slots are zeroed in one memset at the start of a step; a second use within the same step
        # (activation-checkpoint recomputation, a BN module applied twice) must start from zero again
        key = (id(bn), which)
        if key in self._slots_used:
            # through .data: the slot shares one allocation (and therefore one autograd version counter) with the
            # bf16 weights; a tracked in-place write here would invalidate every weight view saved for backward
            slot.tensor.data.zero_()
        self._slots_used.add(key)
        return slot

    def fwd_slot(self, bn) -> _Slot:
        return self._slot(bn, 0)

    def bwd_slot(self, bn) -> _Slot:
        return self._slot(bn, 1)

    def note_bn_step(self, bn):
        if bn.num_batches_tracked is not None:
            self._bn_stepped.append(bn.num_batches_tracked)

    # ------------------------------------------------------------------------------ communication setup
    def _setup_comm(self):
        """Allocate w16 / gradient staging / BN statistics / flags -- in symmetric memory when world > 1."""
        dev = self.device
        n16 = self.total * 2                     # bytes of one bf16 flat buffer
        stats_bytes = 2 * self.stats_len * 4
        flag_bytes = 4096
        layout = {"w16": 0, "stage": n16, "stats": 2 * n16, "flags": 2 * n16 + stats_bytes}
        nbytes = 2 * n16 + stats_bytes + flag_bytes
        self.peer_state = None
        self.comm_state = None
        self.symm_handle = None
        self.syncbn_wait_ns = None
        raw = None
        if self.world > 1 and self.comm_mode == "peer":
            try:
                import torch.distributed._symmetric_memory as symm
                raw = symm.empty(nbytes, dtype=torch.uint8, device=dev)
                raw.zero_()
                torch.cuda.synchronize(dev)
                self.symm_handle = symm.rendezvous(raw, group=dist.group.WORLD)
            except Exception as exc:  # loud, not silent: the run continues on the NCCL baseline path
                logger.warning(f"[b200] symmetric memory unavailable ({type(exc).__name__}: {exc}); "
                               "falling back to B200.COMM=nccl (baseline path, NOT the fused peer kernel)")
                self.comm_mode = "nccl"
                self.symm_handle = None
                raw = None
        if raw is None:
            raw = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self._raw = raw
        self.flat_w16 = raw[layout["w16"]:layout["w16"] + n16].view(torch.bfloat16)
        self.flat_stage = raw[layout["stage"]:layout["stage"] + n16].view(torch.bfloat16)
        self.stats_buf = raw[layout["stats"]:layout["stats"] + stats_bytes].view(torch.float32)
        self.stats_sym_base = 0  # offsets handed to kernels are relative to the per-rank statistics base pointer
        self._local_sync = torch.zeros(16, dtype=torch.int32, device=dev)
        if self.sync_bn and self.comm_mode != "peer":
            raise RuntimeError("MODEL.SYNCBN on the native engine needs the peer-memory path (B200.COMM=peer); "
                               "use B200.ENGINE=torch for SyncBN over NCCL")
        if self.symm_handle is not None:
            h = self.symm_handle
            ptrs = [int(p) for p in h.buffer_ptrs]
            mc = int(h.multicast_ptr) if getattr(h, "multicast_ptr", 0) else 0
            cs = self.K.CommState()
            cs.world, cs.rank, cs.slot_base = self.world, self.rank, 0
            cs.signal_pads = [p + layout["flags"] for p in ptrs]
            cs.stage = [p + layout["stage"] for p in ptrs]
            cs.w16 = [p + layout["w16"] for p in ptrs]
            cs.mc_stage = mc + layout["stage"] if mc else 0
            cs.mc_w16 = mc + layout["w16"] if mc else 0
            cs.local_counter = self._local_sync[0:].data_ptr()
            cs.local_release = self._local_sync[4:].data_ptr()
            # exchange / barrier counters live on the device (kernels read and advance them), so that a captured CUDA
            # graph of the step can be replayed: [0] SyncBN exchanges, [4] gradient all-reduce barriers
            self._epochs = torch.zeros(8, dtype=torch.int32, device=dev)
            cs.epoch_dev = self._epochs[4:].data_ptr()
            self.comm_state = cs
            ps = self.K.PeerState()
            ps.world, ps.rank, ps.slot_base = self.world, self.rank, 64   # flag slots 64.. belong to SyncBN
            ps.signal_pads = [p + layout["flags"] for p in ptrs]
            ps.sym_bufs = [p + layout["stats"] for p in ptrs]
            ps.ticket = self._local_sync[8:].data_ptr()
            ps.epoch_dev = self._epochs[0:].data_ptr()
            # SyncBN: one designated CTA per BN launch reduces the layer's statistics over all ranks (in the switch
            # when the buffer has a multicast mapping) into `reduced` and releases `ready`; the rest read it locally
            max_c = max([m.num_features for m in self.bn_offsets] + [8])
            self._bn_reduced = torch.zeros(2 * max_c, dtype=torch.float32, device=dev)
            self.syncbn_wait_ns = torch.zeros(1, dtype=torch.int64, device=dev)
            ps.mc_stats = mc + layout["stats"] if mc else 0
            ps.reduced = self._bn_reduced.data_ptr()
            ps.ready = self._local_sync[12:].data_ptr()
            ps.wait_ns = self.syncbn_wait_ns.data_ptr()
            self.peer_state = ps
            self.has_multicast = bool(mc)
            dist.barrier()
            if self.rank == 0:
                logger.info(f"[b200] peer-memory comm ready: world={self.world} multicast={'yes' if mc else 'no'} "
                            f"symmetric bytes={nbytes}")

    def _plan_buckets(self, cap_bytes: int):
        """Contiguous flat ranges in reverse parameter order (the order gradients become ready); first bucket 1 MiB.
        The 1-D region is cut into one-shot buckets (see ``_build_flat_storage``)."""
        self.buckets: List[_Bucket] = []
        self.bucket_of: Dict[nn.Parameter, _Bucket] = {}

        def plan(region, first_cap, cap, force_one_shot):
            cur: List[nn.Parameter] = []
            cur_bytes = 0
            limit = first_cap

            def close():
                nonlocal cur, cur_bytes, limit
                if not cur:
                    return
                lo = self.index[cur[-1]][0]
                last_off, last_n = self.index[cur[0]]
                hi = (last_off + last_n + _ALIGN - 1) // _ALIGN * _ALIGN
                b = _Bucket(off=lo, n=hi - lo, params=list(cur),
                            one_shot=force_one_shot or (hi - lo) * 2 <= _ONE_SHOT_BYTES)
                for q in cur:
                    self.bucket_of[q] = b
                self.buckets.append(b)
                cur, cur_bytes, limit = [], 0, cap

            for p in reversed(region):
                nbytes = (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN * 4
                if cur and cur_bytes + nbytes > limit:
                    close()
                cur.append(p)
                cur_bytes += nbytes
            close()

        plan(self._region_big, 1 << 20, cap_bytes, False)
        # one-shot buckets are bounded by the bf16 staging footprint (_ONE_SHOT_BYTES), i.e. 2x that in fp32 bytes
        plan(self._region_1d, 2 * _ONE_SHOT_BYTES, 2 * _ONE_SHOT_BYTES, True)
        self._reset_pending()

    def _reset_pending(self):
        for b in self.buckets:
            b.pending = sum(1 for p in b.params if p.requires_grad)

    def _sync_initial_state(self):
        if self.world > 1:
            dist.broadcast(self.flat_master, src=0)
            for buf in self.module.buffers():
                dist.broadcast(buf, src=0)
        self.refresh_compute_weights()

    @torch.no_grad()
    def sync_buffers(self):
        """Rank 0's buffers (BN running statistics) win -- what reference DDP's ``broadcast_buffers`` does on the
        first eval forward after training (trainer.py:134).  Called by ``validate`` so that the logged accuracy is
        the accuracy of the checkpoint rank 0 writes; with SyncBN the statistics are identical already and the
        broadcast is a no-op in effect."""
        if self.world == 1:
            return
        bufs = [b for b in self.module.buffers() if b.is_floating_point() and b.numel()]
        if not bufs:
            return
        flat = torch.cat([b.reshape(-1).float() for b in bufs])
        dist.broadcast(flat, src=0)
        off = 0
        for b in bufs:
            b.copy_(flat[off:off + b.numel()].view(b.shape))
            off += b.numel()

    def refresh_compute_weights(self):
        self.K.cast_bf16(self.flat_master, self.flat_w16)
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier()

    def on_weights_loaded(self):
        self.refresh_compute_weights()

    # ------------------------------------------------------------------------------ optimizer plumbing
    def make_optimizer(self, **hparams):
        self.optimizer = FusedSGD(self, **hparams)
        return self.optimizer

    def _owned_mask_apply(self, flat):
        """Zero the regions of ``flat`` this rank does not own (see sync_masters)."""
        if self.comm_mode != "peer":
            return
        for b in self.buckets:
            if b.one_shot:
                if self.rank != 0:
                    flat[b.off:b.off + b.n].zero_()
                continue
            n8 = b.n // 8
            per = (n8 + self.world - 1) // self.world
            lo = min(per * self.rank, n8) * 8
            hi = min(lo // 8 + per, n8) * 8
            flat[b.off:b.off + lo].zero_()
            flat[b.off + hi:b.off + b.n].zero_()

    def sync_masters(self):
        """Make every rank's fp32 master weights and momentum complete (two-shot buckets keep only the owner's
        shard up to date).  Collective; called before checkpointing."""
        if self.world == 1 or self.comm_mode != "peer":
            return
        torch.cuda.synchronize(self.device)
        for flat in (self.flat_master, self.flat_mom):
            self._owned_mask_apply(flat)
            dist.all_reduce(flat)

    def mark_ready(self, p):
        b = self.bucket_of.get(p)
        if b is None:
            return
        b.pending -= 1
        if b.pending == 0 and self.world > 1 and self._in_train_step and not self.debug_skip_comm:
            self._launch_bucket(b)

    def _launch_bucket(self, b: _Bucket):
        lr, mom, damp, wd, nest, first = self.optimizer.hyper()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            if self.comm_mode == "peer":
                grid = 8 if b.one_shot else 120   # <= #SMs: every CTA must become resident for the grid barrier
                self.K.allreduce_sgd(self.comm_state, self.flat_master, self.flat_mom, self.flat_grad, b.off, b.n,
                                     lr, mom, damp, wd, nest, first, b.one_shot, grid)
            else:  # NCCL baseline: library all-reduce, then the local fused update
                g = self.flat_grad[b.off:b.off + b.n]
                dist.all_reduce(g)
                self.K.sgd_local(self.flat_master, self.flat_mom, self.flat_grad, self.flat_w16, b.off, b.n,
                                 lr, mom, damp, wd, nest, first, 1.0 / self.world, True)

    _in_train_step = False
    debug_skip_comm = False   # measurement aid: update locally without any gradient exchange (ranks diverge!)
    _deferred = None

    @property
    def defer_wgrad(self) -> bool:
        """SyncBN across GPUs: weight-gradient GEMMs are launched between the two passes of the preceding BN backward."""
        return self.sync_bn and self.world > 1

    def defer(self, fn):
        self.flush_deferred()
        self._deferred = fn

    def flush_deferred(self):
        fn, self._deferred = self._deferred, None
        if fn is not None:
            fn()

    # ------------------------------------------------------------------------------ steps
    def forward(self, x):
        with runtime.native_scope(self):
            return self.module(x)

    def _begin_step(self):
        self._step_parity ^= 1
        base = self._step_parity * self.stats_len
        self.stats_buf[base:base + self.stats_len].zero_()
        self._bn_stepped.clear()
        self.last_sink.clear()
        self._slots_used.clear()

    def train_step(self, inputs, targets, optimizer, topk: int):
        assert optimizer is self.optimizer, "the native engine steps its own FusedSGD (utils.construct_optimizer)"
        if self.cuda_graph and self.device.type == "cuda" and self.module.training and self.comm_mode in ("local", "peer"):
            return self._graphed_step(inputs, targets, optimizer, topk)
        return self._eager_step(inputs, targets, optimizer, topk)

    # ------------------------------------------------------------------------------ CUDA-graph replay
    _GRAPH_WARMUP_STEPS = 3     # eager steps before the first capture (lazy kernel attributes, scratch buffers, TMA maps)

    def _graphed_step(self, inputs, targets, optimizer, topk: int):
        """The whole step -- ~330 kernel launches for ResNet-50, 400-480 for EfficientNet / RegNetY, issued from ~110
        autograd nodes -- costs 8-9 ms of Python per step, which is the wall for the reference's own per-GPU batches of
        32-64.  After a few eager steps the step is captured once per (input shape, hyper-parameters) and replayed:
        inputs are copied into static buffers, outputs (loss, hit counts) are static tensors.  A new learning rate (once
        per epoch) re-captures.  Everything the step does on the host besides launching kernels happens here.

        Multi-GPU steps are captured too -- the gradient-exchange kernels on the side stream join the capture through
        the events that order them, and the SyncBN / all-reduce exchange counters live in device memory (the kernels
        advance them), so nothing in the kernel arguments changes from step to step.  The statistics buffers alternate
        between two halves from step to step (a peer may still be reading the previous step's sums), hence one graph per
        parity."""
        parity = self._step_parity ^ 1
        key = (tuple(inputs.shape), inputs.dtype, tuple(targets.shape), optimizer.hyper(), int(topk), parity,
               bool(self.sync_bn), bool(self.debug_skip_comm))
        entry = self._graphs.get(key)
        if entry is None:
            if self._eager_steps < self._GRAPH_WARMUP_STEPS or not optimizer.has_momentum_state:
                self._eager_steps += 1
                return self._eager_step(inputs, targets, optimizer, topk)
            entry = self._capture(key, inputs, targets, optimizer, topk)
        graph, sx, sy, outs = entry
        sx.copy_(inputs, non_blocking=True)
        sy.copy_(targets, non_blocking=True)
        graph.replay()
        self._step_parity = parity
        optimizer.steps += 1
        self.graph_replays += 1
        return outs

    def _capture(self, key, inputs, targets, optimizer, topk):
        if len(self._graphs) >= 6:           # (two parities per input kind / learning rate) keep the cache small
            self._graphs.pop(next(iter(self._graphs)))
        sx, sy = torch.empty_like(inputs), torch.empty_like(targets)
        sx.copy_(inputs)
        sy.copy_(targets)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        if self._graph_pool is None:
            self._graph_pool = torch.cuda.graph_pool_handle()
        counter = getattr(self.K, "count", None)
        steps_before = optimizer.steps
        parity_before = self._step_parity
        with torch.cuda.graph(graph, pool=self._graph_pool):
            loss, hits1, hitsk = self._eager_step(sx, sy, optimizer, topk)
        optimizer.steps = steps_before       # nothing ran during capture; replay() does the step
        self._step_parity = parity_before
        if counter is not None:
            self.graph_launches_per_step = self.K.count - counter
        entry = (graph, sx, sy, (loss, hits1, hitsk))
        self._graphs[key] = entry
        if self.rank == 0:
            logger.info(f"[b200] captured the training step in a CUDA graph (inputs {tuple(inputs.shape)} {inputs.dtype}, "
                        f"lr {key[3][0]:g}; {len(self._graphs)} graph(s) cached)")
        return entry

    def _eager_step(self, inputs, targets, optimizer, topk: int):
        self._begin_step()
        self._reset_pending()
        self._in_train_step = True
        try:
            with runtime.native_scope(self):
                logits = self.module(inputs)
                loss, hits1, hitsk = self.ops.cross_entropy_topk(logits, targets, topk)
            loss.backward()
            self.flush_deferred()
        finally:
            self._deferred = None
            self._in_train_step = False
        if self._bn_stepped:
            torch._foreach_add_(self._bn_stepped, 1)
        lr, mom, damp, wd, nest, first = optimizer.hyper()
        if self.world == 1 or self.debug_skip_comm:
            if self.trainable_total:
                self.K.sgd_local(self.flat_master, self.flat_mom, self.flat_grad, self.flat_w16, 0,
                                 self.trainable_total, lr, mom, damp, wd, nest, first, 1.0, True)
        else:
            for b in self.buckets:  # parameters that produced no gradient this step still take part
                if b.pending > 0:
                    b.pending = 0
                    self._launch_bucket(b)
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        optimizer.has_momentum_state = True
        optimizer.steps += 1
        return loss.detach(), hits1, hitsk

    @torch.no_grad()
    def eval_step(self, inputs, targets, topk: int):
        with runtime.native_scope(self):
            logits = self.module(inputs)
            return self.ops.cross_entropy_topk(logits, targets, topk)
