"""Functional CPU restatement of the reference U-Net (gelslim_depth/models/unet.py).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference builds the network out of
``torch.nn`` modules; this file restates the same arithmetic as one flat function over a
``state_dict`` so that every intermediate tensor can be tapped, in fp32 or fp64, without
instantiating the reference (which does not exist on the GPU box).

All ``file:line`` citations are relative to /root/reference/.
"""
from __future__ import annotations

import hashlib
from typing import Dict, List

import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm2d default, unet.py:12,15
BN_MOMENTUM = 0.1   # nn.BatchNorm2d default


def _bn_relu(x, sd, prefix, training, stats_out=None):
    """BatchNorm2d + ReLU, unet.py:12-13 / 15-16.

    eval: (x - running_mean) / sqrt(running_var + eps) * gamma + beta.
    train: batch statistics over (N,H,W) with *biased* variance for the normalisation and the
    *unbiased* variance for the running-stat update (momentum 0.1) -- PyTorch semantics.
    """
    g, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if stats_out is not None:
            n = x.numel() // x.shape[1]
            stats_out[prefix] = (mean.detach().clone(), (var * n / max(n - 1, 1)).detach().clone())
    else:
        mean, var = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    inv = torch.rsqrt(var.to(x.dtype) + BN_EPS)
    y = (x - mean.to(x.dtype)[None, :, None, None]) * (inv * g.to(x.dtype))[None, :, None, None] \
        + b.to(x.dtype)[None, :, None, None]
    return torch.relu(y)


def _double_conv(x, sd, prefix, training, taps, stats_out):
    """DoubleConv, unet.py:7-20: [conv3x3 pad 1 no bias -> BN -> ReLU] x 2."""
    for conv_i, bn_i in ((0, 1), (3, 4)):
        w = sd[f"{prefix}.double_conv.{conv_i}.weight"].to(x.dtype)
        x = F.conv2d(x, w, bias=None, stride=1, padding=1)
        if taps is not None:
            taps[f"{prefix}.double_conv.{conv_i}"] = x
        x = _bn_relu(x, sd, f"{prefix}.double_conv.{bn_i}", training, stats_out)
        if taps is not None:
            taps[f"{prefix}.double_conv.{bn_i + 1}"] = x
    return x


def count_levels(sd) -> int:
    n = 0
    while f"down.{n}.maxpool_conv.1.double_conv.0.weight" in sd:
        n += 1
    return n


def unet_forward_with_taps(sd: Dict[str, torch.Tensor], x: torch.Tensor, training: bool = False,
                           dtype=torch.float32, want_taps: bool = True):
    """UNet.forward, unet.py:79-88.  Returns (output, taps, batch_stats).

    taps: every conv output (pre-BN, key '<prefix>.double_conv.{0,3}'), every post-ReLU
    activation ('...{2,5}'), every transposed-conv output ('up.i.up'), the padded+concatenated
    tensor ('up.i.cat') and 'outc'.
    """
    taps = {} if want_taps else None
    stats = {} if training else None
    x = x.to(dtype)
    skips: List[torch.Tensor] = [_double_conv(x, sd, "inc", training, taps, stats)]       # unet.py:80
    depth = count_levels(sd)
    for i in range(depth):                                                               # unet.py:82-83
        p = F.max_pool2d(skips[-1], 2)                                                   # unet.py:26 (floor mode)
        if taps is not None:
            taps[f"down.{i}.pool"] = p
        skips.append(_double_conv(p, sd, f"down.{i}.maxpool_conv.1", training, taps, stats))
    y = skips[-1]
    for i in range(depth):                                                               # unet.py:85-86
        skip = skips[-2 - i]
        w = sd[f"up.{i}.up.weight"].to(dtype)
        b = sd[f"up.{i}.up.bias"].to(dtype)
        k = w.shape[-1]
        u = F.conv_transpose2d(y, w, b, stride=k)                                        # unet.py:36,41
        if taps is not None:
            taps[f"up.{i}.up"] = u
        dy, dx = skip.shape[2] - u.shape[2], skip.shape[3] - u.shape[3]
        u = F.pad(u, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])                     # unet.py:43-47
        cat = torch.cat([skip, u], dim=1)                                                # unet.py:48
        if taps is not None:
            taps[f"up.{i}.cat"] = cat
        y = _double_conv(cat, sd, f"up.{i}.conv", training, taps, stats)
    out = F.conv2d(y, sd["outc.conv.weight"].to(dtype), sd["outc.conv.bias"].to(dtype))  # unet.py:54
    if taps is not None:
        taps["outc"] = out
    return out, taps, stats


def unet_forward(sd, x, training: bool = False, dtype=torch.float32):
    return unet_forward_with_taps(sd, x, training, dtype, want_taps=False)[0]


# --------------------------------------------------------------------------------------
# Synthetic checkpoints
# --------------------------------------------------------------------------------------

def conditioned_state_dict(sd: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Re-draw a state_dict so that activations stay O(1) through all 23 layers.

    A random-init eval-mode U-Net is numerically degenerate (SURVEY.md §4 note 1: the output
    is outc.bias +- 1e-9 with the trainer's N(0, 0.01) init), which makes any max-abs parity
    check vacuous.  Here conv weights get He-normal init, BN gamma ~ U(0.5, 1.5),
    beta ~ U(-0.3, 0.3), running_mean ~ N(0, 0.2), running_var ~ U(0.5, 2).
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            out[k] = v.clone()
        elif k.endswith("running_mean"):
            out[k] = 0.2 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            out[k] = 0.5 + 1.5 * torch.rand(v.shape, generator=g)
        elif v.dim() == 4:       # conv / convT / outc weights
            if ".up.weight" in k:
                fan_in = v.shape[0]                      # each output pixel sees C_in taps once
            else:
                fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            out[k] = torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif k.endswith(".weight"):   # BN gamma
            out[k] = 0.5 + torch.rand(v.shape, generator=g)
        else:                          # BN beta, convT / outc bias
            out[k] = 0.6 * torch.rand(v.shape, generator=g) - 0.3
    return out


def random_init_state_dict(n_channels: int, n_classes: int, dims=(64, 128, 256, 512, 1024), seed: int = 0) -> Dict[str, torch.Tensor]:
    """The state_dict `torch.manual_seed(seed); UNet(n_channels, n_classes, dims)` yields (unet.py:60-77), built from
    plain torch.nn layers created in the reference's construction order (DoubleConv: conv, BN, conv, BN; Down: the same
    behind a MaxPool; Up: ConvTranspose2d then DoubleConv; OutConv last), so the default-init RNG stream and the key
    order are the reference's.  Lets the CPU baseline draw its weights without importing the product package."""
    import torch.nn as nn
    torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def double(prefix, cin, cout):
        for ci, bi, (a, b) in ((0, 1, (cin, cout)), (3, 4, (cout, cout))):
            conv, bn = nn.Conv2d(a, b, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(b)
            sd[f"{prefix}.double_conv.{ci}.weight"] = conv.weight.detach()
            for k, v in bn.state_dict().items():
                sd[f"{prefix}.double_conv.{bi}.{k}"] = v.detach()

    double("inc", n_channels, dims[0])
    for i, (lo, hi) in enumerate(zip(dims[:-1], dims[1:])):
        double(f"down.{i}.maxpool_conv.1", lo, hi)
    for i, (hi, lo) in enumerate(zip(dims[:0:-1], dims[-2::-1])):
        up = nn.ConvTranspose2d(hi, hi // 2, kernel_size=2, stride=2)
        sd[f"up.{i}.up.weight"], sd[f"up.{i}.up.bias"] = up.weight.detach(), up.bias.detach()
        double(f"up.{i}.conv", hi, lo)
    outc = nn.Conv2d(dims[0], n_classes, kernel_size=1)
    sd["outc.conv.weight"], sd["outc.conv.bias"] = outc.weight.detach(), outc.bias.detach()
    return sd


def trainer_init_state_dict(sd: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """train_unet.py:248-250: every parameter whose name contains 'weight' (conv, convT, outc
    AND BatchNorm gamma) is re-drawn N(0, 0.01^2); biases and buffers keep their values."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if "weight" in k and v.is_floating_point():
            out[k] = 0.01 * torch.randn(v.shape, generator=g)
        else:
            out[k] = v.clone()
    return out


def state_dict_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()
