"""Quantisation-aware CPU restatement of the bf16 training step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Same arithmetic as oracle/train_oracle.py (the reference's
train-mode forward + autograd backward, train_utils/train_unet.py:346-374) evaluated in fp32, but every tensor the B200
path stores in bf16 between kernels is rounded to bf16 at exactly that point: GEMM weights, the network input, the
pre-BatchNorm conv outputs (stored centred on the running mean), the post-ReLU activations, the transposed-conv
outputs, and -- in backward -- the activation gradients that pass between kernels: every dgrad output (the gradient of
each conv / transposed-conv INPUT, per consumer), the BatchNorm-backward outputs and the max-pool-backward sums; the last
unit's activation gradient is NOT rounded (it never leaves the fused OutConv / BatchNorm backward kernel).  Batch statistics come from the
un-rounded fp32 conv output, like the GPU epilogue that accumulates them from the fp32 accumulators.

Why it exists: on synthetic random networks with train-mode BatchNorm the gradient is extremely sensitive to 0.2 %
perturbations (rounding only the WEIGHTS to bf16 changes first-layer gradients by ~25 % in a 13-layer net), so the fp32
reference is not a usable pointwise oracle for bf16 gradients.  This restatement is: the GPU path must agree with it
closely, and its own distance to the fp32 reference is the stated inherent bf16 bound.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .unet_oracle import count_levels, BN_EPS


class _RoundBoth(torch.autograd.Function):
    """bf16 storage of a tensor in forward and of its gradient in backward"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


class _RoundBwd(torch.autograd.Function):
    """identity in forward; the gradient is stored in bf16: the dgrad kernels write every conv / transposed-conv INPUT
    gradient as a bf16 tensor (one per consumer, before any summation over consumers)"""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


class _RoundFwd(torch.autograd.Function):
    """bf16 operand (weights, input frames); the gradient w.r.t. it stays fp32 (wgrad accumulates in fp32)"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


def _unit(x, w, gamma, beta, running_mean, last=False):
    z = F.conv2d(_RoundBwd.apply(x), _RoundFwd.apply(w), padding=1)
    mean = z.mean(dim=(0, 2, 3))
    var = z.var(dim=(0, 2, 3), unbiased=False)
    rm = running_mean[None, :, None, None]
    zq = _RoundBoth.apply(z - rm) + rm                     # stored centred on the running mean
    inv = torch.rsqrt(var + BN_EPS)
    a = torch.relu((zq - mean[None, :, None, None]) * (inv * gamma)[None, :, None, None] + beta[None, :, None, None])
    # stored bf16; its gradient (sum over consumers, e.g. max-pool backward + skip connection) is stored bf16 too -- except
    # for the network's last unit, whose activation gradient never leaves the fused OutConv / BatchNorm backward kernel
    return _RoundFwd.apply(a) if last else _RoundBoth.apply(a)


def loss_and_grads_bf16(sd, x, target):
    """-> (loss, {param name: grad}, output) of one train-mode step with bf16 storage points."""
    P = {k: v.detach().clone().float().requires_grad_(True) for k, v in sd.items()
         if v.is_floating_point() and "running" not in k}
    depth = count_levels(sd)

    def dc(t, pre, last=False):
        for ci, bi in ((0, 1), (3, 4)):
            t = _unit(t, P[f"{pre}.double_conv.{ci}.weight"], P[f"{pre}.double_conv.{bi}.weight"],
                      P[f"{pre}.double_conv.{bi}.bias"], sd[f"{pre}.double_conv.{bi}.running_mean"].float(), last=last and ci == 3)
        return t

    t = _RoundFwd.apply(x.float())
    skips = [dc(t, "inc")]
    for i in range(depth):
        skips.append(dc(F.max_pool2d(skips[-1], 2), f"down.{i}.maxpool_conv.1"))
    y = skips[-1]
    for i in range(depth):
        sk = skips[-2 - i]
        u = _RoundBoth.apply(F.conv_transpose2d(_RoundBwd.apply(y), _RoundFwd.apply(P[f"up.{i}.up.weight"]), P[f"up.{i}.up.bias"], stride=2))
        dy_, dx_ = sk.shape[2] - u.shape[2], sk.shape[3] - u.shape[3]
        u = F.pad(u, [dx_ // 2, dx_ - dx_ // 2, dy_ // 2, dy_ - dy_ // 2])
        y = dc(torch.cat([sk, u], 1), f"up.{i}.conv", last=(i == depth - 1))
    out = F.conv2d(y, P["outc.conv.weight"], P["outc.conv.bias"])
    loss = torch.mean((out - target.float()) ** 2)
    grads = torch.autograd.grad(loss, list(P.values()))
    return loss.detach(), dict(zip(P.keys(), grads)), out.detach()
