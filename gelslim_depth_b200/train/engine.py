"""Training step of the U-Net on B200, host side.

Mirrors the loop body of train_utils/train_unet.py:346-377:
    optimizer.zero_grad(); output = unet(x=input); loss = MSE(output, target); loss.backward();
    optimizer.step(); ema.update()
The step itself -- which kernels run, in which order, on which stream, over which buffers -- is sequenced inside
libgsd_b200.so (csrc/train_plan.h: gsd_train_forward / gsd_backward / gsd_adam_ema_step / gsd_train_step).  This module
only owns the memory (torch tensors: flat parameter / gradient / moment arenas and the plan's workspace), hands the
library its pointers, and supplies the one host-side policy of the data-parallel path: which gradients form a bucket and
what to do when a bucket is complete (an NCCL all-reduce on a communication stream).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import torch
import torch.nn as nn

from .. import _lib
from .._lib import lib, check
from . import ops

BF16 = torch.bfloat16


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _bn_modules(net):
    return [m for m in net.modules() if isinstance(m, nn.BatchNorm2d)]


class TrainPlan:
    """One gsd_train_plan + the torch tensor backing its workspace."""

    def __init__(self, net, batch, height, width, device):
        g = _lib.Geometry()
        g.batch, g.in_channels, g.height, g.width, g.n_classes = batch, net.n_channels, height, width, net.n_classes
        g.n_dims = len(net.layer_dimensions)
        for i, d in enumerate(net.layer_dimensions):
            g.dims[i] = int(d)
        # 'fp32' (UNet.set_precision): the FFMA parity path -- same step, fp32 activations, ~1/50 of the speed (csrc/train_plan_f32.h)
        g.dtype = _lib.DTYPE_FP32 if getattr(net, "precision", "bf16") == "fp32" else _lib.DTYPE_BF16
        g.mode = _lib.MODE_TRAIN
        self.precision = getattr(net, "precision", "bf16")
        self.shape = (batch, height, width)
        self.device = device
        self.handle = C.c_void_p()
        check(lib.gsd_train_plan_create(C.byref(self.handle), C.byref(g), device.index or 0), "gsd_train_plan_create")
        self.workspace = torch.empty(lib.gsd_train_plan_workspace_bytes(self.handle), dtype=torch.uint8, device=device)
        self.n_params = lib.gsd_train_plan_num_params(self.handle)
        self.n_bn = lib.gsd_train_plan_num_bn(self.handle)
        self._bound = None

    @property
    def launches(self) -> int:
        return lib.gsd_train_plan_launches(self.handle)

    def bind(self, params, grads, bns):
        """params / grads: tensors in net.parameters() order; bns: BatchNorm2d modules in module order.  Re-binds only
        when a pointer changed."""
        key = tuple(t.data_ptr() for t in params) + tuple(t.data_ptr() for t in grads) + \
            tuple(b.data_ptr() for m in bns for b in (m.running_mean, m.running_var, m.num_batches_tracked))
        if key == self._bound:
            return
        if len(params) != self.n_params or len(grads) != self.n_params or len(bns) != self.n_bn:
            raise ValueError(f"expected {self.n_params} parameters and {self.n_bn} BatchNorm layers")
        for t in list(params) + list(grads) + [b for m in bns for b in (m.running_mean, m.running_var)]:
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("parameters, gradients and BatchNorm statistics must be contiguous fp32 tensors on the plan's device")
        pa = (C.c_void_p * self.n_params)(*[t.data_ptr() for t in params])
        ga = (C.c_void_p * self.n_params)(*[t.data_ptr() for t in grads])
        ba = (C.c_void_p * (2 * self.n_bn))(*[b.data_ptr() for m in bns for b in (m.running_mean, m.running_var)])
        na = (C.c_void_p * self.n_bn)(*[m.num_batches_tracked.data_ptr() for m in bns])
        check(lib.gsd_train_plan_bind(self.handle, pa, ga, ba, na, _p(self.workspace)), "gsd_train_plan_bind")
        self._bound = key

    def set_buckets(self, buckets, bucket_of_index):
        n = len(buckets)
        if n == 0:
            check(lib.gsd_train_plan_set_buckets(self.handle, 0, None, None, None), "gsd_train_plan_set_buckets")
            return
        bo = (C.c_int * self.n_params)(*bucket_of_index)
        lo = (C.c_longlong * n)(*[b["lo"] for b in buckets])
        hi = (C.c_longlong * n)(*[b["hi"] for b in buckets])
        check(lib.gsd_train_plan_set_buckets(self.handle, n, bo, lo, hi), "gsd_train_plan_set_buckets")

    def forward(self, x, y):
        check(lib.gsd_train_forward(self.handle, _p(x), _p(y), ops._st(self.device)), "gsd_train_forward")

    def backward(self, dy, cb=None):
        check(lib.gsd_backward(self.handle, _p(dy), ops._st(self.device), cb or _lib.NULL_CB, None), "gsd_backward")

    def train_step(self, x, target, loss, opt=None, cb=None):
        check(lib.gsd_train_step(self.handle, _p(x), _p(target), _p(loss), C.byref(opt) if opt is not None else None,
                                 ops._st(self.device), cb or _lib.NULL_CB, None), "gsd_train_step")

    def __del__(self):
        try:
            if self.handle:
                lib.gsd_train_plan_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


# ----------------------------------------------------------------------------------------------------------------------
# autograd bridge: `output = unet(x=...)` in .train() mode with stock torch losses / optimizers around it
# ----------------------------------------------------------------------------------------------------------------------
class _BridgeState:
    """per-module state of the autograd bridge: one plan per input shape, one flat gradient buffer"""

    def __init__(self, net):
        params = list(net.parameters())
        dev = params[0].device
        self.flat_g = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in params:
            self.views.append(self.flat_g[off:off + p.numel()].view(p.shape))
            off += p.numel()
        self.plans = {}

    def plan_for(self, net, x):
        key = (x.shape[0], x.shape[2], x.shape[3], x.device, getattr(net, "precision", "bf16"))
        plan = self.plans.get(key)
        if plan is None:
            if len(self.plans) >= 2:
                self.plans.clear()
            plan = self.plans[key] = TrainPlan(net, x.shape[0], x.shape[2], x.shape[3], x.device)
        plan.bind([p.detach() for p in net.parameters()], self.views, _bn_modules(net))
        return plan


class _TrainFn(torch.autograd.Function):
    """`output = unet(x=...)` in .train() mode (train_unet.py:347); `loss.backward()` (train_unet.py:374) lands here."""

    @staticmethod
    def forward(fctx, net, x, *params):
        st = getattr(net, "_train_bridge", None)
        if st is None or st.flat_g.device != x.device:
            st = net._train_bridge = _BridgeState(net)
        plan = st.plan_for(net, x)
        y = torch.empty(x.shape[0], net.n_classes, x.shape[2], x.shape[3], dtype=torch.float32, device=x.device)
        plan.forward(x, y)
        fctx.plan, fctx.state = plan, st
        return y

    @staticmethod
    def backward(fctx, dy):
        fctx.plan.backward(dy.contiguous().float())
        grads = fctx.state.flat_g.clone()           # autograd may keep what we return as .grad: never hand out the live buffer
        out, off = [], 0
        for v in fctx.state.views:
            out.append(grads[off:off + v.numel()].view(v.shape))
            off += v.numel()
        return (None, None, *out)


def unet_train_forward(net, x):
    return _TrainFn.apply(net, x, *list(net.parameters()))


# ----------------------------------------------------------------------------------------------------------------------
# data-parallel gradient reduction (host policy)
# ----------------------------------------------------------------------------------------------------------------------
class BucketReducer:
    """What happens when the library reports a complete gradient bucket (gsd_bucket_cb): `launch(lo, hi)` all-reduces the
    arena range.  On the GPU `launch` orders the communication stream after both compute streams first (FusedTrainer);
    the world_size-2 gloo test (tests/test_ddp_cpu.py) drives this same class from the library's own dry-run backward
    with a plain `dist.all_reduce`.  Also checks the protocol: every bucket exactly once per step."""

    def __init__(self, n_buckets, launch):
        self.n, self.launch = n_buckets, launch
        self.fired = []

    def begin_step(self):
        self.fired = []

    def on_bucket(self, bucket, lo, hi):
        if bucket in self.fired or not 0 <= bucket < self.n:
            raise RuntimeError(f"gradient bucket {bucket} reported twice or out of range")
        self.fired.append(bucket)
        self.launch(lo, hi)

    def all_fired(self) -> bool:
        return sorted(self.fired) == list(range(self.n))


class FusedTrainer:
    """The whole loop body of train_unet.py:346-377 as ONE library call per step (gsd_train_step): parameters, gradients,
    Adam moments and the EMA shadow live in flat fp32 arenas (one kernel updates all 64 tensors); data-parallel replicas
    all-reduce the gradient arena over NCCL in buckets overlapped with backward (per-replica BatchNorm statistics, like
    stock DistributedDataParallel, whose construction-time broadcast of rank 0's model is reproduced too)."""

    def __init__(self, net, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-6, ema_decay=0.995, process_group=None,
                 bucket_bytes=25 << 20, distributed=None, use_graph=False, overlap_wgrad=True, first_bucket_bytes=4 << 20,
                 tail_bucket_bytes=2 << 20):
        self.net = net
        self.lr, self.betas, self.eps, self.wd, self.ema_decay = lr, betas, eps, weight_decay, ema_decay
        params = list(net.parameters())
        n = sum(p.numel() for p in params)
        n_pad = (n + 3) // 4 * 4
        dev = params[0].device
        self.device = dev
        self.flat_p = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        off = 0
        self.views, self.index, self.grad_views = [], {}, []
        for p in params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)          # parameters now alias the arena
            self.grad_views.append(self.flat_g[off:off + k].view_as(p))
            self.views.append((off, k))
            self.index[p] = (off, k)
            off += k
        self.counter = torch.zeros(2, dtype=torch.int64, device=dev)   # (Adam steps, EMA updates) so far, advanced on device
        self.use_graph, self._graph, self._warm = use_graph, None, 0
        self.pg = process_group
        if distributed is None:
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.world = torch.distributed.get_world_size(process_group) if distributed else 1
        self.comm_stream = torch.cuda.Stream(device=dev) if self.world > 1 else None
        if self.world > 1:
            # stock DistributedDataParallel broadcasts rank 0's parameters and buffers at construction: replicas built
            # from different RNG states / checkpoints must start from ONE model, or they average gradients of different
            # networks for ever after
            broadcast_module_state(net, self.flat_p, process_group)
        self.shadow = self.flat_p.clone()                         # torch_ema: shadow = [p.clone()] (after the broadcast)
        self.buckets, self.bucket_of = plan_buckets([(i, *self.views[i]) for i in range(len(params))], bucket_bytes,
                                                    first_bucket_bytes=first_bucket_bytes, tail_bucket_bytes=tail_bucket_bytes)
        self.overlap_wgrad = overlap_wgrad
        self.plan = None
        self._loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.reducer = BucketReducer(len(self.buckets), self._allreduce_range) if self.world > 1 else None
        self._cb_streams = (None, None)
        # the ctypes thunk lives as long as the trainer but must not keep it alive (a cycle would park the captured CUDA
        # graph -- with its NCCL nodes -- until interpreter shutdown, after the process group is gone)
        ref = weakref.ref(self)

        def _thunk(user, bucket, lo, hi, main_stream, side_stream):
            me = ref()
            if me is not None:
                me._on_bucket(user, bucket, lo, hi, main_stream, side_stream)
        self._cb = _lib.BUCKET_CB(_thunk)
        self._opt = _lib.OptimizerState()
        o = self._opt
        o.params, o.grads, o.m, o.v, o.ema = (self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                              self.shadow.data_ptr())
        o.n, o.counter = n_pad, self.counter.data_ptr()
        o.hp.lr, o.hp.beta1, o.hp.beta2, o.hp.eps, o.hp.weight_decay = lr, betas[0], betas[1], eps, weight_decay
        o.hp.ema_decay, o.hp.grad_scale = ema_decay, 1.0 / self.world

    # ------------------------------------------------------------------ data-parallel callback
    def _on_bucket(self, _user, bucket, lo, hi, main_stream, side_stream):
        """gsd_bucket_cb: called by the library, on this thread, while it enqueues the step"""
        if self.reducer is None or not self._reduce:
            return
        if bucket < 0:          # backward fully enqueued, the optimizer kernel is next: it must see the reduced gradients
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
            return
        self._cb_streams = (main_stream, side_stream)
        self.reducer.on_bucket(bucket, lo, hi)

    def _allreduce_range(self, lo, hi):
        # A bucket mixes gradients written on the main stream (BatchNorm, bias) and on the weight-gradient side stream:
        # the all-reduce waits for BOTH, whichever stream happened to launch the bucket's last kernel.
        for ptr in self._cb_streams:
            if ptr:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.ExternalStream(ptr, device=self.device))
                self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            torch.distributed.all_reduce(self.flat_g[lo:hi], group=self.pg)

    # ------------------------------------------------------------------ plan
    def _plan_for(self, x):
        shape = (x.shape[0], x.shape[2], x.shape[3])
        if self.plan is None or self.plan.shape != shape or self.plan.precision != getattr(self.net, "precision", "bf16"):
            if self._graph is not None:
                raise RuntimeError("use_graph=True: the input shape is fixed once the step has been captured")
            if not self.overlap_wgrad:
                os.environ["GSD_NO_WGRAD_OVERLAP"] = "1"
            self.plan = TrainPlan(self.net, x.shape[0], x.shape[2], x.shape[3], x.device)
            if not self.overlap_wgrad:
                os.environ.pop("GSD_NO_WGRAD_OVERLAP", None)
            index_of = [self.bucket_of[i] for i in range(len(self.views))]
            self.plan.set_buckets(self.buckets if self.world > 1 else [], index_of)
        self.plan.bind([p.detach() for p in self.net.parameters()], self.grad_views, _bn_modules(self.net))
        return self.plan

    def close(self):
        """release the captured graph, the plan and its workspace now (before a process group is destroyed)"""
        self._graph = None
        self.plan = None

    def average_parameters(self):
        """`with trainer.average_parameters():` == torch_ema's context manager (train_unet.py:389,428,480): the EMA
        shadow is swapped into the parameters for validation / checkpointing and the live weights are restored on exit
        (two arena copies; BatchNorm running statistics stay live, exactly like the reference)."""
        trainer = self

        class _Ctx:
            def __enter__(self_inner):
                trainer._stash = trainer.flat_p.clone()
                trainer.flat_p.copy_(trainer.shadow)
                return trainer

            def __exit__(self_inner, *exc):
                trainer.flat_p.copy_(trainer._stash)
                trainer._stash = None
                return False

        return _Ctx()

    def step(self, x, target) -> torch.Tensor:
        """one training step; returns the loss as a 1-element device tensor (no host sync).
        A NaN loss is NOT replaced by a constant (train_unet.py:371-373 does that and would then crash in
        backward): the step runs and the NaN is visible to the caller.
        use_graph=True: after two eager warm-up steps the whole step (incl. the NCCL all-reduces) is captured once
        into a CUDA graph and replayed; shapes must then stay fixed."""
        if not self.use_graph:
            return self._step_impl(x, target).clone()
        if self._graph is None:
            if self._warm < 2:
                self._warm += 1
                return self._step_impl(x, target).clone()
            self._x = x.contiguous().float().clone()
            self._t = target.contiguous().float().clone()
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._step_impl(self._x, self._t)
            # capture records but does not execute: replay once so that this call is a real step too
        self._x.copy_(x)
        self._t.copy_(target)
        self._graph.replay()
        return self._loss.clone()

    def _step_impl(self, x, target, optimize=True, reduce=True) -> torch.Tensor:
        x, target = x.contiguous().float(), target.contiguous().float()
        plan = self._plan_for(x)
        self._reduce = reduce
        if self.reducer is not None:
            self.reducer.begin_step()
        plan.train_step(x, target, self._loss, self._opt if optimize else None, self._cb if (self.world > 1 and reduce) else None)
        if self.reducer is not None and reduce and not self.reducer.all_fired():
            raise RuntimeError("data-parallel step: not every gradient bucket was reported")
        return self._loss

    def backward_only(self, x, target, reduce=True) -> torch.Tensor:
        """forward + loss + backward WITHOUT the optimizer update; returns a copy of the flat gradient arena (summed over
        ranks when `reduce`, this rank's own gradient otherwise).  Verification aid for the data-parallel path
        (bench.py / tests: reduced gradient == mean of per-rank gradients); BatchNorm running statistics do advance."""
        self._step_impl(x, target, optimize=False, reduce=reduce)
        if self.world > 1 and reduce:
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)
        return self.flat_g.clone()

    def param_checksum(self) -> torch.Tensor:
        """(sum p, sum p^2, sum shadow) in fp64 -- replicas in sync hold bit-identical values"""
        p, sh = self.flat_p.double(), self.shadow.double()
        return torch.stack([p.sum(), (p * p).sum(), sh.sum()])


def plan_buckets(entries, bucket_bytes, first_bucket_bytes=None, tail_bucket_bytes=None):
    """entries: [(key, offset, numel)] in parameter (= arena) order.  Gradients appear in REVERSE order during backward,
    so buckets are contiguous arena ranges grown from the end.  -> ([{lo, hi, params}], {key: bucket index})

    first_bucket_bytes: size at which the FIRST bucket closes (small: the first all-reduce starts right after the head /
    last decoder block instead of half-way through backward).  tail_bucket_bytes: the parameters whose gradients arrive
    LAST (the start of the arena: inc.*, down.0.*) form their own bucket of at most about this size, so the all-reduce
    that cannot overlap anything is short."""
    rev = list(reversed(entries))
    n_tail = 0
    if tail_bucket_bytes:
        acc = 0
        for key, off, k in entries:                      # arena order == reverse arrival order
            if acc + 4 * k > tail_bucket_bytes and n_tail > 0:
                break
            acc += 4 * k
            n_tail += 1
            if acc >= tail_bucket_bytes:
                break
        if n_tail >= len(entries):
            n_tail = 0
    body, tail = (rev[:len(rev) - n_tail], rev[len(rev) - n_tail:]) if n_tail else (rev, [])
    buckets, bucket_of = [], {}
    cur = None
    for key, off, k in body:
        if cur is None:
            cur = {"lo": off, "hi": off + k, "params": []}
        cur["lo"] = off
        cur["params"].append(key)
        limit = first_bucket_bytes if (first_bucket_bytes and not buckets) else bucket_bytes
        if (cur["hi"] - cur["lo"]) * 4 >= limit:
            buckets.append(cur)
            cur = None
    if cur is not None:
        buckets.append(cur)
    if tail:
        buckets.append({"lo": tail[-1][1], "hi": tail[0][1] + tail[0][2], "params": [key for key, _, _ in tail]})
    for i, b in enumerate(buckets):
        for key in b["params"]:
            bucket_of[key] = i
    return buckets, bucket_of


def broadcast_module_state(net, flat_p, process_group=None, src=0):
    """rank `src`'s parameters (one flat arena) and BatchNorm buffers -> every rank (DistributedDataParallel's
    construction-time broadcast, torch/nn/parallel/distributed.py `_sync_module_states`)."""
    dist = torch.distributed
    dist.broadcast(flat_p, src, group=process_group)
    bufs = [b for b in net.buffers()]
    if not bufs:
        return
    fl = [b for b in bufs if b.is_floating_point()]
    it = [b for b in bufs if not b.is_floating_point()]
    for group in (fl, it):
        if not group:
            continue
        flat = torch.cat([b.detach().reshape(-1) for b in group])
        dist.broadcast(flat, src, group=process_group)
        off = 0
        for b in group:
            b.data.copy_(flat[off:off + b.numel()].view_as(b))
            off += b.numel()
