"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The fixtures pin oracle/ (tests/test_oracle_golden.py) and are the reference-generated targets
of the CUDA parity tests (tests/test_gpu_parity.py).  Weights of the full-width nets are NOT
stored (124 MB); they are regenerated from the recorded seed by
oracle.conditioned_state_dict / trainer_init_state_dict and checked against the recorded
sha256 digest.
"""
import os
import re
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, "/root/reference")

import oracle  # noqa: E402
from gelslim_depth.models.unet import UNet  # noqa: E402
from gelslim_depth.processing_utils import image_utils, normalization_utils  # noqa: E402
from gelslim_depth.processing_utils.complete_prediction import predict_depth_from_RGB  # noqa: E402

torch.set_num_threads(8)
FULL = [64, 128, 256, 512, 1024]
SMALL = [4, 8, 16, 32, 64]
TAP_KEYS = ["inc.double_conv.0", "inc.double_conv.5", "down.0.pool", "down.1.maxpool_conv.1.double_conv.5",
            "down.3.maxpool_conv.1.double_conv.5", "up.0.up", "up.0.cat", "up.0.conv.double_conv.5",
            "up.3.up", "up.3.conv.double_conv.5", "outc"]


def shipped_config(size):
    """Attribute set of gelslim_depth/config/config_unet_bigdata.py (:23-43), under BOTH
    spellings of the image-normalisation attributes (complete_prediction.py:6 reads tactile_*)."""
    c = types.SimpleNamespace()
    c.input_tactile_image_size = size
    c.interp_method = "area"
    c.norm_scale = 0.9
    c.image_normalization_method = c.tactile_normalization_method = "0_255_to_0_1"
    c.image_normalization_parameters = c.tactile_normalization_parameters = None
    c.depth_normalization_method = "min_max_to_0_-1"
    c.depth_normalization_parameters = (-1.9180814027786255, 0.0)
    return c


def hooks_for(model):
    """Forward hooks that capture the same tensors oracle.unet_forward_with_taps names."""
    store = {}

    def mk(name):
        def h(_m, _i, o):
            store[name] = o.detach().clone()
        return h

    def mk_pre_clone(name):   # ReLU is inplace -> the conv output must be cloned before BN/ReLU
        def h(_m, _i, o):
            store[name] = o.detach().clone()
        return h

    for name, mod in model.named_modules():
        if isinstance(mod, (torch.nn.Conv2d,)) and "double_conv" in name:
            mod.register_forward_hook(mk_pre_clone(name))
        elif isinstance(mod, torch.nn.ReLU):
            mod.register_forward_hook(mk(name))
        elif isinstance(mod, torch.nn.ConvTranspose2d):
            mod.register_forward_hook(mk(name))
        elif isinstance(mod, torch.nn.MaxPool2d):
            mod.register_forward_hook(mk(name.replace("maxpool_conv.0", "pool")))
        elif name == "outc":
            mod.register_forward_hook(mk("outc"))
        elif re.fullmatch(r"up\.\d+\.conv", name):        # input of the decoder DoubleConv == cat([skip, up])
            def pre(_m, inp, _name=name.replace(".conv", ".cat")):
                store[_name] = inp[0].detach().clone()
            mod.register_forward_pre_hook(pre)
    return store


def summarise(t):
    f = t.flatten()
    idx = torch.linspace(0, f.numel() - 1, 64).long()
    return dict(shape=list(t.shape), mean=float(t.mean()), std=float(t.std()), absmax=float(t.abs().max()),
                sample_idx=idx, sample=f[idx].clone())


def make_small():
    """Small-width nets: full state_dict stored; eval + train forward, grads, 6 optimiser steps."""
    out = {}
    for tag, (cin, ncls, h, w) in {"c3": (3, 1, 40, 53), "c6": (6, 2, 32, 43)}.items():
        torch.manual_seed(11)
        net = UNet(cin, ncls, layer_dimensions=SMALL)
        sd = oracle.conditioned_state_dict(net.state_dict(), seed=5)
        net.load_state_dict(sd)
        g = torch.Generator().manual_seed(3)
        x = torch.rand(2, cin, h, w, generator=g)
        net.eval()
        store = hooks_for(net)
        with torch.no_grad():
            y_eval = net(x=x)
        taps_eval = {k: store[k] for k in TAP_KEYS}
        # one train-mode forward/backward
        net2 = UNet(cin, ncls, layer_dimensions=SMALL)
        net2.load_state_dict(sd)
        net2.train()
        tgt = -0.9 * torch.rand(2, ncls, h, w, generator=g)
        y_train = net2(x=x)
        loss = torch.mean((y_train - tgt) ** 2)
        loss.backward()
        grads = {k: p.grad.clone() for k, p in net2.named_parameters()}
        bn_after = {k: v.clone() for k, v in net2.state_dict().items() if "running" in k or "tracked" in k}
        # 6 Adam steps with the trainer's hyper-parameters (train_unet.py:306); EMA is restated
        # in oracle/train_oracle.py because torch_ema is not installed here (parity for EMA is
        # therefore pinned only to its published formula).
        net3 = UNet(cin, ncls, layer_dimensions=SMALL)
        net3.load_state_dict(sd)
        net3.train()
        opt = torch.optim.Adam(net3.parameters(), lr=1e-3, weight_decay=1e-6)
        losses = []
        for _ in range(6):
            opt.zero_grad()
            l = torch.mean((net3(x=x) - tgt) ** 2)
            l.backward()
            opt.step()
            losses.append(float(l.detach()))
        out[tag] = dict(cin=cin, ncls=ncls, dims=SMALL, state_dict=sd, x=x, target=tgt, y_eval=y_eval,
                        taps_eval=taps_eval, y_train=y_train.detach(), loss=float(loss.detach()), grads=grads,
                        bn_after=bn_after, adam_losses=losses,
                        params_after={k: summarise(p.detach()) for k, p in net3.named_parameters()})
    torch.save(out, os.path.join(HERE, "unet_small.pt"))


def make_full():
    """Full-width nets (the dims the CUDA kernels run): weights by seed + digest."""
    out = {}
    cases = {"g2_eval": (6, 2, 2, 48, 59, "conditioned"), "g3_eval": (3, 1, 2, 32, 43, "conditioned"),
             "g1_train": (3, 1, 2, 32, 43, "trainer")}
    for tag, (cin, ncls, b, h, w, init) in cases.items():
        torch.manual_seed(0)
        net = UNet(cin, ncls, layer_dimensions=FULL)
        base_sd = net.state_dict()
        sd = (oracle.conditioned_state_dict if init == "conditioned" else oracle.trainer_init_state_dict)(base_sd, seed=7)
        net.load_state_dict(sd)
        g = torch.Generator().manual_seed(21)
        x = torch.rand(b, cin, h, w, generator=g)
        rec = dict(cin=cin, ncls=ncls, dims=FULL, init=init, init_seed=7, module_seed=0,
                   digest=oracle.state_dict_digest(sd), x=x)
        if init == "conditioned":
            net.eval()
            store = hooks_for(net)
            with torch.no_grad():
                rec["y"] = net(x=x)
            rec["taps"] = {k: summarise(store[k]) for k in TAP_KEYS}
        else:
            net.train()
            tgt = -0.9 * torch.rand(b, ncls, h, w, generator=g)
            opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-6)
            losses = []
            for _ in range(8):
                opt.zero_grad()
                y = net(x=x)
                l = torch.mean((y - tgt) ** 2)
                l.backward()
                if not losses:
                    rec["grad_summ"] = {k: summarise(p.grad) for k, p in net.named_parameters()
                                        if k in ("inc.double_conv.0.weight", "down.3.maxpool_conv.1.double_conv.3.weight",
                                                 "up.0.up.weight", "up.3.conv.double_conv.4.bias", "outc.conv.weight",
                                                 "outc.conv.bias", "down.1.maxpool_conv.1.double_conv.1.weight")}
                    rec["y0"] = y.detach().clone()
                opt.step()
                losses.append(float(l.detach()))
            rec["target"] = tgt
            rec["adam_losses"] = losses
        out[tag] = rec
    torch.save(out, os.path.join(HERE, "unet_full.pt"))


def make_processing():
    """Entry point + helpers (complete_prediction.py:4-10, image_utils.py, normalization_utils.py)."""
    out = {}
    g = torch.Generator().manual_seed(9)
    raw = torch.randint(0, 256, (2, 6, 64, 85), generator=g).float()
    base = torch.randint(0, 256, (1, 6, 64, 85), generator=g).float()
    diff = image_utils.get_difference_image(raw, base)
    out["raw"], out["base"], out["diff"] = raw, base, diff
    fingers = torch.cat([diff[:, 0:3], diff[:, 3:6]], dim=0)              # general_dataset.py:71
    torch.manual_seed(11)
    net = UNet(3, 1, layer_dimensions=SMALL)
    sd = oracle.conditioned_state_dict(net.state_dict(), seed=5)
    net.load_state_dict(sd)
    net.eval()
    cfg = shipped_config((32, 43))
    with torch.no_grad():
        out["depth"] = predict_depth_from_RGB(fingers, net, (64, 85), cfg)
    out["state_dict_digest"] = oracle.state_dict_digest(sd)   # == unet_small.pt['c3']['state_dict']
    # area resampling at the real geometry (320x427 <-> 160x213), on a small deterministic image
    img = torch.rand(1, 1, 320, 427, generator=g)
    down = image_utils.sample_multi_channel_image_to_desired_size(img, (160, 213), "area")
    up = image_utils.sample_multi_channel_image_to_desired_size(down, (320, 427), "area")
    out["area_seed_img_digest"] = float(img.double().sum())
    out["area_down"] = down
    out["area_up_summ"] = summarise(up)
    # every normalisation method
    t = torch.rand(2, 3, 5, 7, generator=g) * 255
    p4 = ([1.0, 2.0, 3.0], [200.0, 210.0, 220.0], [100.0, 110.0, 120.0], [50.0, 60.0, 70.0])
    out["norm_in"] = t
    for m in ("mean_std", "0_255_to_-1_1", "0_255_to_0_1"):
        out["norm_img_" + m] = normalization_utils.normalize_tactile_image(t, m, 0.9, p4)
    # normalization_utils.py:9 evaluates `0.5*(...).tolist()` == float * list -> the reference
    # raises TypeError for 'min_max_to_-1_1' image normalisation; record that as the behaviour.
    try:
        normalization_utils.normalize_tactile_image(t, "min_max_to_-1_1", 0.9, p4)
        out["norm_img_min_max_to_-1_1_raises"] = None
    except Exception as e:  # noqa: BLE001
        out["norm_img_min_max_to_-1_1_raises"] = type(e).__name__
    d = -2.0 * torch.rand(2, 1, 5, 7, generator=g)
    out["depth_in"] = d
    dp = (-1.9180814027786255, 0.0, -0.4, 0.3)
    for m in ("min_max_to_-1_1", "mean_std", "min_max_to_0_1", "min_max_to_0_-1"):
        out["norm_depth_" + m] = normalization_utils.normalize_depth_image(d, m, 0.9, dp)
        out["denorm_depth_" + m] = normalization_utils.denormalize_depth_image(d, m, 0.9, dp)
    torch.save(out, os.path.join(HERE, "processing.pt"))


if __name__ == "__main__":
    make_small()
    make_full()
    make_processing()
    for f in ("unet_small.pt", "unet_full.pt", "processing.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
