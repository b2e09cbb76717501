"""CPU restatement of the training-step body (train_utils/train_unet.py:340-377).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Third-party arithmetic restated here because it is not under /root/reference:
  * torch.optim.Adam (torch, unpinned "v2.0+", README.md:13) with lr=1e-3, betas=(0.9,0.999),
    eps=1e-8, weight_decay=1e-6 *coupled* L2 (train_unet.py:306): published algorithm
        g <- g + wd*p;  m <- b1*m + (1-b1)*g;  v <- b2*v + (1-b2)*g*g
        p <- p - lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
  * torch_ema==0.3 ExponentialMovingAverage(params, decay=0.995) (requirements.txt:6,
    train_unet.py:309,376), use_num_updates=True:
        n <- n+1;  d <- min(decay, (1+n)/(10+n));  shadow <- shadow - (1-d)*(shadow - p)
"""
from __future__ import annotations

from typing import Dict

import torch

from .unet_oracle import unet_forward_with_taps, BN_MOMENTUM


def mse_loss(output, target):
    """train_unet.py:51-52."""
    return torch.mean((output - target) ** 2)


def param_keys(sd):
    """nn.Module.parameters() order == state_dict order minus BN buffers."""
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


class TrainOracle:
    def __init__(self, sd: Dict[str, torch.Tensor], lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=1e-6, ema_decay=0.995, dtype=torch.float32):
        self.dtype = dtype
        self.sd = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        self.keys = param_keys(self.sd)
        self.lr, self.betas, self.eps, self.wd, self.ema_decay = lr, betas, eps, weight_decay, ema_decay
        self.m = {k: torch.zeros_like(self.sd[k]) for k in self.keys}
        self.v = {k: torch.zeros_like(self.sd[k]) for k in self.keys}
        self.shadow = {k: self.sd[k].clone() for k in self.keys}
        self.t = 0
        self.ema_updates = 0
        self.last_grads = None

    def loss_and_grads(self, x, target):
        leaves = {k: self.sd[k].detach().clone().requires_grad_(True) for k in self.keys}
        sd = dict(self.sd)
        sd.update(leaves)
        out, _, stats = unet_forward_with_taps(sd, x, training=True, dtype=self.dtype, want_taps=False)
        loss = mse_loss(out, target.to(self.dtype))                           # train_unet.py:370
        grads = torch.autograd.grad(loss, [leaves[k] for k in self.keys])      # train_unet.py:374
        return loss.detach(), dict(zip(self.keys, grads)), stats, out.detach()

    def step(self, x, target) -> float:
        loss, grads, stats, _ = self.loss_and_grads(x, target)
        self.last_grads = grads
        # BatchNorm running statistics (train-mode forward side effect)
        for prefix, (mean, var_unbiased) in stats.items():
            self.sd[prefix + ".running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean)
            self.sd[prefix + ".running_var"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var_unbiased)
            self.sd[prefix + ".num_batches_tracked"] += 1
        self.apply_update(grads)
        return float(loss)

    def apply_update(self, grads):
        """optimizer.step(); ema.update() (train_unet.py:375-376) for given gradients {key: tensor}"""
        # Adam, coupled L2 (train_unet.py:306,375)
        self.t += 1
        b1, b2 = self.betas
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        for k in self.keys:
            p = self.sd[k]
            g = grads[k] + self.wd * p
            self.m[k].mul_(b1).add_((1 - b1) * g)
            self.v[k].mul_(b2).add_((1 - b2) * g * g)
            denom = self.v[k].sqrt() / (bc2 ** 0.5) + self.eps
            p.sub_((self.lr / bc1) * self.m[k] / denom)
        # EMA (train_unet.py:309,376)
        self.ema_updates += 1
        d = min(self.ema_decay, (1 + self.ema_updates) / (10 + self.ema_updates))
        for k in self.keys:
            self.shadow[k].sub_((1 - d) * (self.shadow[k] - self.sd[k]))
