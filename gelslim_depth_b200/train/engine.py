"""Training step of the U-Net on B200: train-mode forward, backward and (optionally) the fused Adam+EMA update.

Mirrors the loop body of train_utils/train_unet.py:346-377:
    optimizer.zero_grad(); output = unet(x=input); loss = MSE(output, target); loss.backward();
    optimizer.step(); ema.update()
Every FLOP runs in libgsd_b200.so: conv forward / dgrad on the halo-resident tcgen05 kernel, wgrad as a tcgen05 GEMM
over pixels, BatchNorm statistics in the conv epilogue, everything else as fused memory-bound kernels
(csrc/train_ops.cuh).  This module only sequences the launches and owns no arithmetic.
"""
from __future__ import annotations

import contextlib
from typing import List

import torch
import torch.nn as nn

from . import ops

BF16 = torch.bfloat16


class _Unit:
    """saved tensors of one conv3x3 -> BatchNorm -> ReLU unit"""
    __slots__ = ("src0", "src1", "off", "z", "a", "mean", "rstd", "scale", "shift", "pooled", "conv", "bn", "first")


def _blocks(net):
    enc = [net.inc.double_conv] + [d.maxpool_conv[1].double_conv for d in net.down]
    dec = [(u.up, u.conv.double_conv) for u in net.up]
    return enc, dec


class PackedTrainWeights:
    """bf16 GEMM operands of every layer for forward and for dgrad.  The operand buffers and a device table of
    (parameter pointer, mode, shape, output pointer) records are built once; `repack()` refreshes all of them with ONE
    kernel launch (gsd_op_pack_weights_batched) after each optimizer step."""

    def __init__(self, net):
        from .._lib import PackItem
        enc, dec = _blocks(net)
        self.fwd, self.dgrad = {}, {}
        from .._lib import lib
        items = []          # (mode, param, O, I, Ipad, key, wants dgrad operand)
        for bi, seq in enumerate(enc):
            for ci in (0, 3):
                w = seq[ci].weight
                O, I = w.shape[:2]
                first = bi == 0 and ci == 0
                items.append((0, w, O, I, 16 if first else I, id(seq[ci]), not first))
        for up, seq in dec:
            I, O = up.weight.shape[:2]
            items.append((2, up.weight, O, I, I, id(up), False))
            items.append((3, up.weight, O, I, I, id(up), False))
            for ci in (0, 3):
                cw = seq[ci].weight
                items.append((0, cw, cw.shape[0], cw.shape[1], cw.shape[1], id(seq[ci]), True))
        dev = items[0][1].device
        al = lambda n: (n + 7) // 8 * 8                                  # every operand 16-byte aligned
        elems = sum(al(ops.pack_out_elems(m, O, I, ip)) + (al(ops.pack_out_elems(1, O, I)) if dg else 0)
                    for (m, _, O, I, ip, _, dg) in items)
        self.arena = torch.empty(elems, dtype=BF16, device=dev)
        table = (PackItem * len(items))()
        self._ptrs = []
        cur, units = 0, 0
        for k, (mode, w, O, I, ipad, key, dg) in enumerate(items):
            n = ops.pack_out_elems(mode, O, I, ipad)
            out = self.arena[cur:cur + n]
            cur += al(n)
            (self.dgrad if mode == 3 else self.fwd)[key] = out
            table[k].w, table[k].out, table[k].out_dgrad = w.data_ptr(), out.data_ptr(), None
            if dg:
                n2 = ops.pack_out_elems(1, O, I)
                self.dgrad[key] = self.arena[cur:cur + n2]
                table[k].out_dgrad = self.dgrad[key].data_ptr()
                cur += al(n2)
            table[k].mode, table[k].O, table[k].I, table[k].Ipad, table[k].start = mode, O, I, ipad, units
            units += lib.gsd_pack_item_units(mode, O, I, ipad)
            self._ptrs.append((w, w.data_ptr()))
        self.n_items, self.total = len(items), units
        self.table = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(dev)
        self.repack()

    def valid_for(self, net) -> bool:
        return all(w.data_ptr() == ptr for w, ptr in self._ptrs)

    def repack(self):
        ops.pack_weights_batched(self.table, self.n_items, self.total, self.arena.device)


class StepEnv:
    """Per-step scratch policy.  Default: fresh zeroed tensors and one negate launch per BatchNorm (autograd bridge);
    FusedTrainer substitutes persistent arenas (one memset + one negate per step, no per-layer fills)."""

    def zeros(self, n, dev):
        return torch.zeros(n, dtype=torch.float32, device=dev)

    def neg_center(self, bn):
        return ops.negate(bn.running_mean)

    def dwk(self, conv, cin_total):
        return None

    # weight-gradient kernels may run on a side stream, concurrently with the memory-bound BatchNorm-backward passes of
    # the next unit (see _ArenaEnv); the default runs everything on the current stream
    def mark(self):
        """token for 'everything launched so far on the current stream'"""
        return None

    def side(self, token):
        """context manager: kernels launched inside run after `token`, possibly on another stream"""
        return contextlib.nullcontext()

    def keep(self, *tensors):
        """tensors read by side-stream kernels must outlive the Python scope that created them"""

    def join(self):
        """the current stream waits for everything issued under side()"""


_DEFAULT_ENV = StepEnv()


def _unit_forward(conv: nn.Conv2d, bn: nn.BatchNorm2d, pw: PackedTrainWeights, src0, src1=None, off=(0, 0), pool=False, first=False,
                  env: StepEnv = _DEFAULT_ENV, apply=True):
    B, H, W, _ = src0.shape
    cout = conv.out_channels
    stats = env.zeros(2 * cout, src0.device)
    # z is stored centred on the running mean (bf16 then rounds relative to the fluctuation of z, not to its mean);
    # the batch statistics are taken from the raw fp32 accumulators in the conv epilogue.
    neg_center = env.neg_center(bn)
    z = ops.conv(src0, pw.fwd[id(conv)], cout, 9, src1=src1, off=off, shift=neg_center, stats=stats)
    # batch statistics -> (scale, shift, mean, rstd); running stats (momentum 0.1) and num_batches_tracked updated in place
    scale, shift, mean, rstd = ops.bn_finalize(stats, B * H * W, bn, neg_center)
    # apply=False (last unit): the consumer (the 1x1 head) applies BatchNorm + ReLU itself, `a` is never stored
    a, pooled = ops.bn_relu_apply(z, scale, shift, pool=pool) if apply else (None, None)
    u = _Unit()
    u.src0, u.src1, u.off, u.z, u.a, u.mean, u.rstd, u.pooled, u.conv, u.bn, u.first = src0, src1, off, z, a, mean, rstd, pooled, conv, bn, first
    u.scale, u.shift = scale, shift
    return u


def train_forward(net, x: torch.Tensor, pw: PackedTrainWeights, env: StepEnv = _DEFAULT_ENV):
    """-> (y fp32 NCHW, saved context).  BatchNorm uses batch statistics and updates its running buffers."""
    enc, dec = _blocks(net)
    depth = len(enc) - 1
    ctx = {"enc": [], "dec": []}
    cur = ops.prologue(x.contiguous().float())
    ctx["in16"] = cur
    for l, seq in enumerate(enc):
        u1 = _unit_forward(seq[0], seq[1], pw, cur, first=(l == 0), env=env)
        u2 = _unit_forward(seq[3], seq[4], pw, u1.a, pool=(l < depth), env=env)
        ctx["enc"].append((u1, u2))
        cur = u2.pooled if l < depth else u2.a
    y_prev = ctx["enc"][depth][1].a
    for i, (up, seq) in enumerate(dec):
        l = depth - 1 - i
        skip = ctx["enc"][l][1].a
        cup = up.out_channels
        dev = y_prev.device
        u = ops.conv(y_prev, pw.fwd[id(up)], cup, ntaps=1, groups=4, shift=up.bias.detach().repeat(4))
        off = ((skip.shape[1] - u.shape[1]) // 2, (skip.shape[2] - u.shape[2]) // 2)           # F.pad left/top (unet.py:46-47)
        u1 = _unit_forward(seq[0], seq[1], pw, skip, src1=u, off=off, env=env)
        u2 = _unit_forward(seq[3], seq[4], pw, u1.a, env=env, apply=(i < len(dec) - 1))
        ctx["dec"].append((up, y_prev, u, off, u1, u2))
        y_prev = u2.a
    w_head = net.outc.conv.weight.detach().reshape(net.n_classes, -1)
    last = ctx["dec"][-1][5] if dec else None
    if last is not None:
        # OutConv reads relu(BatchNorm(z)) of the last unit directly from z (unet.py:17,54-57 in one pass)
        y = ops.bn_relu_head_fwd(last.z, last.scale, last.shift, w_head, net.outc.conv.bias.detach())
    else:
        y = ops.head_fwd(y_prev, w_head, net.outc.conv.bias.detach())
    ctx["a_last"] = y_prev
    return y, ctx


class GradSink:
    """Where parameter gradients land.  Default: fresh tensors (autograd bridge).  FusedTrainer overrides `dest` with
    views of its flat gradient arena and `done` with the bucketed all-reduce trigger."""

    def __init__(self):
        self.grads = {}

    def dest(self, p):
        t = torch.empty_like(p)
        self.grads[p] = t
        return t

    def put(self, p, value):
        self.dest(p).copy_(value.reshape(p.shape))
        self.done(p)

    def done(self, p):
        pass


def _unit_backward(u: _Unit, da, pw: PackedTrainWeights, grads: "GradSink", need_dx: bool, split: int = 0,
                   env: StepEnv = _DEFAULT_ENV, head=None):
    """backward of conv -> BN -> ReLU.  Returns the input gradient(s) (None if not needed).
    split > 0: the conv input was the virtual concat [skip | up]; returns (dskip, dup_full).
    head = (dy, w_head, dw_head, db_head): this is the last unit, `da` is None and the OutConv backward is fused in."""
    B, H, W, Cn = u.z.shape
    if head is not None:
        dy, w_head, dwh, dbh = head
        dz, sums = ops.head_bn_bwd(u.z, dy, w_head, u.scale, u.shift, u.mean, u.rstd, u.bn.weight.detach(), dwh, dbh,
                                   sums=env.zeros(2 * Cn, dy.device))
    else:
        dz, sums = ops.bn_bwd(da, u.scale, u.shift, u.z, u.mean, u.rstd, u.bn.weight.detach(), B * H * W,
                              sums=env.zeros(2 * Cn, da.device))
    grads.put(u.bn.bias, sums[:Cn])
    grads.put(u.bn.weight, sums[Cn:])

    def weight_grad():
        gw = grads.dest(u.conv.weight)
        if u.first:
            ops.wgrad_first(u.src0, dz, u.conv.in_channels, gw, dwk=env.dwk(u.conv, 16))
        else:
            ops.wgrad3x3(u.src0, dz, gw, x1=u.src1, off=u.off, dwk=env.dwk(u.conv, u.conv.in_channels))
        grads.done(u.conv.weight)

    # wgrad depends only on dz: it is issued (possibly on a side stream) AFTER the dgrad launch, so that dgrad -- which the
    # critical path waits for -- gets the SMs first, and wgrad then overlaps the next unit's HBM-bound BatchNorm passes
    token = env.mark()
    out = None
    if need_dx:
        wd = pw.dgrad[id(u.conv)]                               # [ci][9][co]
        cin = u.conv.in_channels
        if split:
            rows = split * 9 * Cn
            out = (ops.conv(dz, wd[:rows], split, 9), ops.conv(dz, wd[rows:], cin - split, 9))
        else:
            out = ops.conv(dz, wd, cin, 9)
    with env.side(token):
        weight_grad()
    env.keep(dz)
    return out


def train_backward(net, ctx, dy: torch.Tensor, pw: PackedTrainWeights, grads: "GradSink" = None,
                   env: StepEnv = _DEFAULT_ENV) -> List[torch.Tensor]:
    """dy: gradient of the loss w.r.t. the network output (fp32 NCHW).  Gradients go to `grads` (a GradSink) in
    reverse parameter order; returns them in net.parameters() order when the default sink is used."""
    enc, dec = _blocks(net)
    depth = len(enc) - 1
    own = grads is None
    grads = GradSink() if own else grads
    dev = dy.device
    w_head = net.outc.conv.weight.detach().reshape(net.n_classes, -1)
    dbh = grads.dest(net.outc.conv.bias)
    dwh = grads.dest(net.outc.conv.weight)
    dbh.zero_()
    dwh.zero_()
    dy = dy.contiguous().float()
    fused_head = depth > 0                                   # the last unit's activation was never stored
    da = None if fused_head else ops.head_bwd(ctx["a_last"], dy, w_head, dwh, dbh)
    if not fused_head:
        grads.done(net.outc.conv.bias)
        grads.done(net.outc.conv.weight)
    dskips = [None] * (depth + 1)
    # ---- decoder, last block first
    for i in range(depth - 1, -1, -1):
        up, y_prev, u, off, u1, u2 = ctx["dec"][i]
        last = fused_head and i == depth - 1
        da1 = _unit_backward(u2, da, pw, grads, need_dx=True, env=env, head=(dy, w_head, dwh, dbh) if last else None)
        if last:
            grads.done(net.outc.conv.bias)
            grads.done(net.outc.conv.weight)
        cskip = u1.src0.shape[-1]
        dskip, dup = _unit_backward(u1, da1, pw, grads, need_dx=True, split=cskip, env=env)
        l = depth - 1 - i
        dskips[l] = dskip
        # transposed conv: only the (2hs x 2ws) window of dup at `off` is its output gradient (the rest is F.pad)
        hs, ws = y_prev.shape[1], y_prev.shape[2]
        if dup.shape[1] != 2 * hs or dup.shape[2] != 2 * ws:
            dup[:, : off[0]] = 0
            dup[:, off[0] + 2 * hs:] = 0
            dup[:, :, : off[1]] = 0
            dup[:, :, off[1] + 2 * ws:] = 0
        token = env.mark()
        da = ops.convt_dgrad(dup, off, pw.dgrad[id(up)], up.in_channels, hs, ws)
        sums_up = env.zeros(2 * dup.shape[-1], dev)
        with env.side(token):
            grads.put(up.bias, ops.channel_sum(dup, sums=sums_up))
            gw = grads.dest(up.weight)
            ops.convt_wgrad(y_prev, dup, off, gw)
            grads.done(up.weight)
        env.keep(dup)
    # ---- encoder, bottom up: `da` is now the gradient of enc[depth]'s output
    for l in range(depth, -1, -1):
        u1, u2 = ctx["enc"][l]
        if l < depth:
            da = ops.maxpool_bwd(u2.a, dpool, dskips[l])      # noqa: F821  (dpool from level l+1) + skip-connection gradient
        da1 = _unit_backward(u2, da, pw, grads, need_dx=True, env=env)
        dpool = _unit_backward(u1, da1, pw, grads, need_dx=(l > 0), env=env)
    env.join()
    return [grads.grads[p] for p in net.parameters()] if own else None


class _TrainFn(torch.autograd.Function):
    """autograd bridge: `output = unet(x=...)` in .train() mode; `loss.backward()` lands here."""

    @staticmethod
    def forward(fctx, net, x, *params):
        pw = getattr(net, "_train_pw", None)
        if pw is None or not pw.valid_for(net):
            pw = net._train_pw = PackedTrainWeights(net)      # operand buffers + pointer table, built once
        else:
            pw.repack()                                       # parameters changed in place (optimizer.step): one launch
        y, ctx = train_forward(net, x, pw)
        fctx.net, fctx.saved, fctx.pw = net, ctx, pw
        return y

    @staticmethod
    def backward(fctx, dy):
        grads = train_backward(fctx.net, fctx.saved, dy, fctx.pw)
        fctx.saved = None
        return (None, None, *grads)


def unet_train_forward(net, x):
    return _TrainFn.apply(net, x, *list(net.parameters()))


class _ArenaEnv(StepEnv):
    """FusedTrainer's scratch: BatchNorm running means alias one flat arena (one negate launch per step gives every
    layer's centring constant), all small zero-initialised buffers (statistics, reduction sums, the loss) are slices of
    one arena cleared by a single memset, and the weight-gradient accumulators persist (the unpack kernel re-zeroes
    them)."""

    def __init__(self, net, overlap_wgrad=True):
        bns = [m for m in net.modules() if isinstance(m, nn.BatchNorm2d)]
        dev = bns[0].running_mean.device
        total = sum(m.num_features for m in bns)
        self.flat_rm = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat_neg = torch.empty_like(self.flat_rm)
        self._neg = {}
        off = 0
        for m in bns:
            c = m.num_features
            self.flat_rm[off:off + c].copy_(m.running_mean)
            m.running_mean.data = self.flat_rm[off:off + c]       # the buffer now aliases the arena (state_dict unchanged)
            self._neg[id(m)] = self.flat_neg[off:off + c]
            off += c
        self.arena = torch.zeros(1 << 18, dtype=torch.float32, device=dev)
        self.cursor = 0
        self._dwk = {}
        self.side_stream = torch.cuda.Stream(device=dev) if overlap_wgrad else None
        self._kept, self._forked = [], False

    def begin_step(self):
        self._kept = []                   # the previous step's side-stream work was joined in train_backward
        self.arena.zero_()
        self.cursor = 0
        ops.negate(self.flat_rm, out=self.flat_neg)

    def mark(self):
        if self.side_stream is None:
            return None
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        return ev

    @contextlib.contextmanager
    def side(self, token):
        if self.side_stream is None:
            yield
            return
        self.side_stream.wait_event(token)
        self._forked = True
        with torch.cuda.stream(self.side_stream):
            yield

    def keep(self, *tensors):
        if self.side_stream is not None:
            self._kept.extend(tensors)

    def join(self):
        if self._forked:
            torch.cuda.current_stream().wait_stream(self.side_stream)
            self._forked = False

    def zeros(self, n, dev):
        n4 = (n + 3) // 4 * 4
        if self.cursor + n4 > self.arena.numel():
            return torch.zeros(n, dtype=torch.float32, device=dev)
        out = self.arena[self.cursor:self.cursor + n]
        self.cursor += n4
        return out

    def neg_center(self, bn):
        return self._neg[id(bn)]

    def dwk(self, conv, cin_total):
        t = self._dwk.get(id(conv))
        if t is None:
            t = self._dwk[id(conv)] = torch.zeros(conv.out_channels, 9, cin_total, dtype=torch.float32, device=conv.weight.device)
        return t


class BucketReducer:
    """Host logic of the overlapped gradient all-reduce: counts down the parameters of each bucket as their gradient
    kernels are LAUNCHED (backward order) and, when a bucket is complete, hands its contiguous arena range to `launch`.
    `launch(lo, hi)` is the only device-specific part: on the GPU it orders the communication stream after BOTH compute
    streams and enqueues the NCCL all-reduce; the world_size-2 gloo test (tests/test_ddp_cpu.py) passes a plain
    `dist.all_reduce`.  One instance per step."""

    def __init__(self, buckets, bucket_of, launch):
        self.buckets, self.bucket_of, self.launch = buckets, bucket_of, launch
        self.pending = [len(b["params"]) for b in buckets]
        self.fired = []

    def done(self, key):
        bi = self.bucket_of[key]
        self.pending[bi] -= 1
        if self.pending[bi] < 0:
            raise RuntimeError("gradient reported twice for one parameter")
        if self.pending[bi] == 0:
            b = self.buckets[bi]
            self.fired.append(bi)
            self.launch(b["lo"], b["hi"])

    def all_fired(self) -> bool:
        return all(n == 0 for n in self.pending)


class _ArenaSink(GradSink):
    """Gradients are written straight into the flat arena; a bucket (contiguous range formed in backward =
    reverse-parameter order) is all-reduced on the communication stream as soon as its last gradient kernel has been
    launched, so NCCL traffic overlaps the remaining dgrad / wgrad kernels."""

    def __init__(self, trainer, reduce=True):
        super().__init__()
        self.t = trainer
        self.reducer = BucketReducer(trainer.buckets, trainer.bucket_of, self._launch) if (trainer.world > 1 and reduce) else None

    def dest(self, p):
        off, k = self.t.index[p]
        return self.t.flat_g[off:off + k].view(p.shape)

    def _launch(self, lo, hi):
        # A bucket mixes gradients written on the main stream (BatchNorm, bias) and on the weight-gradient side stream:
        # the all-reduce must wait for BOTH, whichever stream happened to launch the bucket's last kernel.
        t = self.t
        streams = {id(s): s for s in (t.main_stream_for_step, torch.cuda.current_stream())}
        if t.env.side_stream is not None and t.env._forked:      # forked this step (inside the graph capture, if any)
            streams[id(t.env.side_stream)] = t.env.side_stream
        for st in streams.values():
            ev = torch.cuda.Event()
            ev.record(st)
            t.comm_stream.wait_event(ev)
        with torch.cuda.stream(t.comm_stream):
            torch.distributed.all_reduce(t.flat_g[lo:hi], group=t.pg)

    def done(self, p):
        if self.reducer is not None:
            self.reducer.done(p)


class FusedTrainer:
    """The whole loop body of train_unet.py:346-377 with the loss, Adam (coupled L2) and the torch_ema update fused:
    parameters, gradients, Adam moments and the EMA shadow live in flat fp32 arenas (one kernel updates all 64 tensors);
    data-parallel replicas all-reduce the gradient arena over NCCL in buckets overlapped with backward
    (per-replica BatchNorm statistics, like stock DistributedDataParallel)."""

    def __init__(self, net, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-6, ema_decay=0.995, process_group=None,
                 bucket_bytes=25 << 20, distributed=None, use_graph=False, overlap_wgrad=True, first_bucket_bytes=4 << 20,
                 tail_bucket_bytes=2 << 20):
        self.net = net
        self.lr, self.betas, self.eps, self.wd, self.ema_decay = lr, betas, eps, weight_decay, ema_decay
        params = list(net.parameters())
        n = sum(p.numel() for p in params)
        n_pad = (n + 3) // 4 * 4
        dev = params[0].device
        self.flat_p = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        off = 0
        self.views, self.index = [], {}
        for p in params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)          # parameters now alias the arena
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
        self.main_stream_for_step = torch.cuda.current_stream(dev)
        if self.world > 1:
            # stock DistributedDataParallel broadcasts rank 0's parameters and buffers at construction: replicas built
            # from different RNG states / checkpoints must start from ONE model, or they average gradients of different
            # networks for ever after
            broadcast_module_state(net, self.flat_p, process_group)
        self.shadow = self.flat_p.clone()                         # torch_ema: shadow = [p.clone()] (after the broadcast)
        self.buckets, self.bucket_of = plan_buckets([(p, *self.index[p]) for p in params], bucket_bytes,
                                                    first_bucket_bytes=first_bucket_bytes, tail_bucket_bytes=tail_bucket_bytes)
        self.pw = PackedTrainWeights(net)                          # after aliasing: the table holds arena pointers
        # overlap_wgrad: weight-gradient GEMMs (tensor-bound) run on a side stream concurrently with the next unit's
        # BatchNorm-backward passes (HBM-bound)
        import os
        self.env = _ArenaEnv(net, overlap_wgrad=overlap_wgrad and not os.environ.get("GSD_NO_WGRAD_OVERLAP"))

    def average_parameters(self):
        """`with trainer.average_parameters():` == torch_ema's context manager (train_unet.py:389,428,480): the EMA
        shadow is swapped into the parameters for validation / checkpointing and the live weights are restored on exit
        (two arena copies; BatchNorm running statistics stay live, exactly like the reference)."""
        trainer = self

        class _Ctx:
            def __enter__(self_inner):
                trainer._stash = trainer.flat_p.clone()
                trainer.flat_p.copy_(trainer.shadow)
                trainer.net.invalidate_packed_weights()
                return trainer

            def __exit__(self_inner, *exc):
                trainer.flat_p.copy_(trainer._stash)
                trainer._stash = None
                trainer.net.invalidate_packed_weights()
                return False

        return _Ctx()

    def step(self, x, target) -> torch.Tensor:
        """one training step; returns the loss as a 1-element device tensor (no host sync).
        A NaN loss is NOT replaced by a constant (train_unet.py:371-373 does that and would then crash in
        backward): the step runs and the NaN is visible to the caller.
        use_graph=True: after two eager warm-up steps the whole step (~450 launches incl. the NCCL all-reduces)
        is captured once into a CUDA graph and replayed; shapes must then stay fixed."""
        if not self.use_graph:
            return self._step_impl(x, target).clone()        # the loss lives in the per-step scratch arena
        if self._graph is None:
            if self._warm < 2:
                self._warm += 1
                return self._step_impl(x, target).clone()
            self._x = x.contiguous().float().clone()
            self._t = target.contiguous().float().clone()
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._loss = self._step_impl(self._x, self._t)
            # capture records but does not execute: replay once so that this call is a real step too
        self._x.copy_(x)
        self._t.copy_(target)
        self._graph.replay()
        self.net.invalidate_packed_weights()
        return self._loss.clone()

    def backward_only(self, x, target, reduce=True) -> torch.Tensor:
        """forward + loss + backward WITHOUT the optimizer update; returns a copy of the flat gradient arena (summed over
        ranks when `reduce`, this rank's own gradient otherwise).  Verification aid for the data-parallel path
        (bench.py / tests: reduced gradient == mean of per-rank gradients); BatchNorm running statistics do advance."""
        net, pw, env = self.net, self.pw, self.env
        self.main_stream_for_step = torch.cuda.current_stream()
        pw.repack()
        env.begin_step()
        y, ctx = train_forward(net, x, pw, env=env)
        _, dy = ops.mse(y, target.contiguous().float(), loss=env.zeros(1, y.device))
        train_backward(net, ctx, dy, pw, grads=_ArenaSink(self, reduce=reduce), env=env)
        if self.world > 1 and reduce:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        net.invalidate_packed_weights()
        return self.flat_g.clone()

    def param_checksum(self) -> torch.Tensor:
        """(sum p, sum p^2, sum shadow) in fp64 -- replicas in sync hold bit-identical values"""
        p, sh = self.flat_p.double(), self.shadow.double()
        return torch.stack([p.sum(), (p * p).sum(), sh.sum()])

    def _step_impl(self, x, target) -> torch.Tensor:
        net, pw, env = self.net, self.pw, self.env
        self.main_stream_for_step = torch.cuda.current_stream()
        pw.repack()                                                # every layer's bf16 operands: one launch
        env.begin_step()                                           # one memset + one negate for all layers
        y, ctx = train_forward(net, x, pw, env=env)
        loss, dy = ops.mse(y, target.contiguous().float(), loss=env.zeros(1, y.device))
        train_backward(net, ctx, dy, pw, grads=_ArenaSink(self), env=env)
        scale = 1.0
        if self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            scale = 1.0 / self.world
        ops.adam_ema_dev(self.flat_p, self.flat_g, self.m, self.v, self.shadow, self.lr, self.betas, self.eps, self.wd,
                         self.ema_decay, self.counter, grad_scale=scale)
        net.invalidate_packed_weights()          # the arena kernel wrote through raw pointers
        return loss


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
