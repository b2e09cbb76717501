"""ctypes wrappers of the training-step operators of libgsd_b200.so (include/gsd_b200.h, "training-step
operators").  torch tensors are only the memory; every wrapper is exactly one kernel launch."""
from __future__ import annotations

import ctypes as C

import torch

from .._lib import lib, check

BF16 = torch.bfloat16


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _st(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


_consts = {}


def ones(dev, n=4096):
    key = ("1", dev)
    if key not in _consts:
        _consts[key] = torch.ones(8192, dtype=torch.float32, device=dev)
    return _consts[key][:n]


def zeros(dev, n=4096):
    key = ("0", dev)
    if key not in _consts:
        _consts[key] = torch.zeros(8192, dtype=torch.float32, device=dev)
    return _consts[key][:n]


def prologue(x):
    """(B,C,H,W) fp32 NCHW -> (B,H,W,16) bf16 NHWC (channels zero padded)."""
    B, Cc, H, W = x.shape
    out = torch.empty(B, H, W, 16, dtype=BF16, device=x.device)
    s8 = (C.c_float * 8)(*([1.0] * 8))
    t8 = (C.c_float * 8)(*([0.0] * 8))
    check(lib.gsd_op_prologue_bf16(_p(x), None, 1, 0, B, Cc, H, W, H, W, s8, t8, _p(out), _st(x.device)), "gsd_op_prologue_bf16")
    return out


def conv(src0, w, cout, ntaps=9, groups=1, src1=None, off=(0, 0), scale=None, shift=None, relu=False, pool=False, stats=None):
    B, H, W, C0 = src0.shape
    dev = src0.device
    out = torch.empty((B, H, W, cout) if groups == 1 else (B, 2 * H, 2 * W, cout), dtype=BF16, device=dev)
    pooled = torch.empty(B, H // 2, W // 2, cout, dtype=BF16, device=dev) if pool else None
    C1 = H1 = W1 = 0
    if src1 is not None:
        _, H1, W1, C1 = src1.shape
    check(lib.gsd_op_conv_auto_bf16(_p(src0), C0, _p(src1), C1, H1, W1, off[0], off[1], B, H, W, _p(w), cout, ntaps, groups,
                                    _p(scale), _p(shift), int(relu), _p(out), _p(pooled), _p(stats), dev.index or 0, _st(dev)),
          "gsd_op_conv_auto_bf16")
    return (out, pooled) if pool else out


def negate(v, out=None):
    out = torch.empty_like(v) if out is None else out
    check(lib.gsd_op_negate_f32(_p(v), v.numel(), _p(out), _st(v.device)), "gsd_op_negate_f32")
    return out


def bn_finalize(stats, count, bn, neg_center=None):
    """-> (scale, shift, mean, rstd); also updates running_mean / running_var and num_batches_tracked += 1 in place."""
    Cn = bn.num_features
    dev = stats.device
    out = torch.empty(4, Cn, dtype=torch.float32, device=dev)
    scale, shift, mean, rstd = out[0], out[1], out[2], out[3]
    check(lib.gsd_op_bn_finalize(_p(stats), float(count), _p(bn.weight), _p(bn.bias), _p(bn.running_mean), _p(bn.running_var),
                                 float(bn.momentum), float(bn.eps), Cn, _p(neg_center), _p(scale), _p(shift), _p(mean), _p(rstd),
                                 _p(bn.num_batches_tracked), _st(dev)), "gsd_op_bn_finalize")
    return scale, shift, mean, rstd


def bn_relu_apply(z, scale, shift, pool=False):
    B, H, W, Cn = z.shape
    a = torch.empty_like(z)
    pooled = torch.empty(B, H // 2, W // 2, Cn, dtype=BF16, device=z.device) if pool else None
    check(lib.gsd_op_bn_relu_apply(_p(z), _p(scale), _p(shift), B, H, W, Cn, _p(a), _p(pooled), _st(z.device)), "gsd_op_bn_relu_apply")
    return a, pooled


def mse(y, t, loss=None):
    loss = torch.zeros(1, dtype=torch.float32, device=y.device) if loss is None else loss   # accumulated: must be zero
    dy = torch.empty_like(y)
    check(lib.gsd_op_mse(_p(y), _p(t), y.numel(), _p(loss), _p(dy), _st(y.device)), "gsd_op_mse")
    return loss, dy


def head_fwd(a, w, bias):
    B, H, W, _ = a.shape
    ncls = w.shape[0]
    y = torch.empty(B, ncls, H, W, dtype=torch.float32, device=a.device)
    check(lib.gsd_op_head_fwd(_p(a), _p(w), _p(bias), ncls, B, H, W, _p(y), _st(a.device)), "gsd_op_head_fwd")
    return y


def bn_relu_head_fwd(z, scale, shift, w, bias):
    """y = OutConv(relu(z*scale + shift)) from the raw conv output of the last unit (its activation is never stored)."""
    B, H, W, _ = z.shape
    ncls = w.shape[0]
    y = torch.empty(B, ncls, H, W, dtype=torch.float32, device=z.device)
    check(lib.gsd_op_bn_relu_head_fwd(_p(z), _p(scale), _p(shift), _p(w), _p(bias), ncls, B, H, W, _p(y), _st(z.device)),
          "gsd_op_bn_relu_head_fwd")
    return y


def head_bn_bwd(z, dy, w, scale, shift, mean, rstd, gamma, dw, db, sums=None):
    """backward of OutConv + ReLU + BatchNorm of the last unit -> (dz bf16, sums = [dbeta | dgamma]); dw / db accumulated."""
    B, H, W, Cn = z.shape
    sums = torch.zeros(2 * Cn, dtype=torch.float32, device=z.device) if sums is None else sums
    dz = torch.empty_like(z)
    check(lib.gsd_op_head_bn_bwd(_p(z), _p(dy), _p(w), _p(scale), _p(shift), _p(mean), _p(rstd), _p(gamma), float(B * H * W),
                                 w.shape[0], B, H, W, _p(sums), _p(dw), _p(db), _p(dz), _st(z.device)), "gsd_op_head_bn_bwd")
    return dz, sums


def head_bwd(a, dy, w, dw, db):
    B, H, W, _ = a.shape
    da = torch.empty_like(a)
    check(lib.gsd_op_head_bwd(_p(a), _p(dy), _p(w), w.shape[0], B, H, W, _p(da), _p(dw), _p(db), _st(a.device)), "gsd_op_head_bwd")
    return da


def bn_bwd(da, scale, shift, z, mean, rstd, gamma, count, sums=None):
    """-> dz (bf16), sums = [dbeta | dgamma] (fp32, 2C; a zeroed buffer may be passed in).
    scale/shift: the forward's BN-apply constants (ReLU mask)."""
    Cn = z.shape[-1]
    npix = z.numel() // Cn
    sums = torch.zeros(2 * Cn, dtype=torch.float32, device=z.device) if sums is None else sums
    check(lib.gsd_op_bn_bwd_reduce(_p(da), _p(scale), _p(shift), _p(z), _p(mean), _p(rstd), npix, Cn, _p(sums), _st(z.device)),
          "gsd_op_bn_bwd_reduce")
    dz = torch.empty_like(z)
    check(lib.gsd_op_bn_bwd_apply(_p(da), _p(scale), _p(shift), _p(z), _p(mean), _p(rstd), _p(gamma), _p(sums), float(count), npix, Cn,
                                  _p(dz), _st(z.device)), "gsd_op_bn_bwd_apply")
    return dz, sums


def channel_sum(t, sums=None):
    """per-channel sum of a dense NHWC bf16 tensor -> fp32 [C]."""
    Cn = t.shape[-1]
    sums = torch.zeros(2 * Cn, dtype=torch.float32, device=t.device) if sums is None else sums
    check(lib.gsd_op_bn_bwd_reduce(_p(t), None, None, None, None, None, t.numel() // Cn, Cn, _p(sums), _st(t.device)),
          "gsd_op_bn_bwd_reduce")
    return sums[:Cn]


def maxpool_bwd(a, dpool, dskip):
    B, H, W, Cn = a.shape
    dfull = torch.empty_like(a)
    check(lib.gsd_op_maxpool_bwd(_p(a), _p(dpool), _p(dskip), B, H, W, Cn, _p(dfull), _st(a.device)), "gsd_op_maxpool_bwd")
    return dfull


def pack_weight(mode, w, O, I, Ipad=None):
    Ipad = I if Ipad is None else Ipad
    n = {0: O * 9 * Ipad, 1: O * 9 * I, 2: 4 * O * I, 3: 4 * O * I}[mode]
    out = torch.empty(n, dtype=BF16, device=w.device)
    check(lib.gsd_op_pack_weight(mode, _p(w), O, I, Ipad, _p(out), _st(w.device)), "gsd_op_pack_weight")
    return out


def wgrad3x3(x0, dz, grad_out, x1=None, off=(0, 0), dwk=None):
    """conv weight gradient written into grad_out (O,I,3,3) fp32.  `dwk`: persistent ZEROED [O][9][I] fp32 accumulation
    buffer; the unpack kernel clears it again as it reads, so it is ready for the next step (no per-step fill)."""
    B, H, W, C0 = x0.shape
    cout = dz.shape[-1]
    dev = x0.device
    C1 = H1 = W1 = 0
    if x1 is not None:
        _, H1, W1, C1 = x1.shape
    clear = dwk is not None
    if dwk is None:
        dwk = torch.zeros(cout, 9, C0 + C1, dtype=torch.float32, device=dev)
    check(lib.gsd_op_wgrad3x3_bf16(_p(x0), C0, _p(x1), C1, H1, W1, off[0], off[1], _p(dz), cout, B, H, W, _p(dwk), dev.index or 0,
                                   _st(dev)), "gsd_op_wgrad3x3_bf16")
    check(lib.gsd_op_unpack_wgrad(_p(dwk), cout, C0 + C1, C0 + C1, _p(grad_out), int(clear), _st(dev)), "gsd_op_unpack_wgrad")


def wgrad_first(x16, dz, cin, grad_out, dwk=None):
    B, H, W, _ = x16.shape
    clear = dwk is not None
    if dwk is None:
        dwk = torch.zeros(64, 9, 16, dtype=torch.float32, device=x16.device)
    # tcgen05 GEMM over pixels with the 16-channel padded input as a 32-byte-row (SWIZZLE_32B, N = 3 x 16) operand
    check(lib.gsd_op_wgrad3x3_bf16(_p(x16), 16, None, 0, 0, 0, 0, 0, _p(dz), 64, B, H, W, _p(dwk), x16.device.index or 0,
                                   _st(x16.device)), "gsd_op_wgrad3x3_bf16")
    check(lib.gsd_op_unpack_wgrad(_p(dwk), 64, cin, 16, _p(grad_out), int(clear), _st(x16.device)), "gsd_op_unpack_wgrad")


def pack_out_elems(mode, O, I, Ipad=None):
    Ipad = I if Ipad is None else Ipad
    return {0: O * 9 * Ipad, 1: O * 9 * I, 2: 4 * O * I, 3: 4 * O * I}[mode]


def pack_weights_batched(table_dev, n_items, total, dev):
    """one launch for every layer: table_dev = device uint8 tensor holding n_items gsd_pack_item records"""
    check(lib.gsd_op_pack_weights_batched(_p(table_dev), n_items, total, _st(dev)), "gsd_op_pack_weights_batched")


def convt_dgrad(du_full, off, w_dgrad, cin, hs, ws):
    B, Hf, Wf, Cs = du_full.shape
    dev = du_full.device
    out = torch.empty(B, hs, ws, cin, dtype=BF16, device=dev)
    check(lib.gsd_op_convt_dgrad_bf16(_p(du_full), Cs, Hf, Wf, off[0], off[1], _p(w_dgrad), cin, B, hs, ws, None, None,
                                      _p(out), dev.index or 0, _st(dev)), "gsd_op_convt_dgrad_bf16")
    return out


def convt_wgrad(x_in, du_full, off, grad_out):
    B, hs, ws, cin = x_in.shape
    _, Hf, Wf, cout = du_full.shape
    dev = x_in.device
    grad_out.zero_()
    check(lib.gsd_op_convt_wgrad_bf16(_p(x_in), cin, _p(du_full), cout, Hf, Wf, off[0], off[1], B, hs, ws, _p(grad_out),
                                      dev.index or 0, _st(dev)), "gsd_op_convt_wgrad_bf16")


def adam_ema(p, g, m, v, shadow, lr, betas, eps, wd, step, ema_decay, ema_updates, grad_scale=1.0):
    check(lib.gsd_op_adam_ema(_p(p), _p(g), _p(m), _p(v), _p(shadow), p.numel(), lr, betas[0], betas[1], eps, wd, step, ema_decay,
                              ema_updates, grad_scale, _st(p.device)), "gsd_op_adam_ema")


def adam_ema_dev(p, g, m, v, shadow, lr, betas, eps, wd, ema_decay, counter, grad_scale=1.0):
    """CUDA-graph-replayable Adam+EMA: `counter` is a device int64[2] (steps, EMA updates) advanced on the device."""
    check(lib.gsd_op_adam_ema_dev(_p(p), _p(g), _p(m), _p(v), _p(shadow), p.numel(), lr, betas[0], betas[1], eps, wd, ema_decay,
                                  _p(counter), grad_scale, _st(p.device)), "gsd_op_adam_ema_dev")
