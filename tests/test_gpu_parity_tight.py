"""Parity where round 1 was loose (VERDICT r1 "what's weak" #1, "missing" #4/#5/#7/#8): per-layer taps against the oracle
(bf16 and fp32 mode, small and full 6x320x427 geometry), the fp32 <= 1e-3 mm bound at the north-star geometry, the
optimizer / EMA kernels over the WHOLE arena, every parameter gradient against the bf16 restatement with a stated bound,
and a 200-step loss curve at full geometry against the fp32 restatement run with torch on the same GPU."""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
MM_PER_UNIT = 1.9180814027786255 / 0.9        # config_unet_bigdata.py:42-43: 1 network unit = (max - min) / norm_scale mm


def dev():
    return torch.device("cuda:0")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def make_net(cin, ncls, seed, dims=(64, 128, 256, 512, 1024)):
    from gelslim_depth_b200.models.unet import UNet
    torch.manual_seed(seed)
    net = UNet(cin, ncls, layer_dimensions=list(dims))
    sd = oracle.conditioned_state_dict(net.state_dict(), seed=seed + 1)
    net.load_state_dict(sd)
    return net, sd


@pytest.mark.parametrize("geom", ["small", "full"])
def test_per_layer_taps_vs_oracle_bf16_and_fp32(geom, monkeypatch):
    """SURVEY 4 note 1a: every conv->BN->ReLU output and every transposed-conv output of the GPU path against the
    oracle's taps (fp64 restatement of unet.py:79-88), so that an error is localised to a layer.
    Bounds: fp32 mode rel-L2 <= 1e-5 per layer and <= 1e-3 mm on the depth map (the north-star bound, here ALSO at
    6x320x427); bf16 mode rel-L2 <= 2e-2 per layer (one bf16 rounding per stored tensor, 22 layers deep)."""
    monkeypatch.setenv("GSD_NO_HEAD_FUSION", "1")          # keep the last unit's activation so that it can be tapped
    B, H, W = (2, 48, 59) if geom == "small" else (1, 320, 427)
    net, sd = make_net(6, 2, 31)
    x = torch.rand(B, 6, H, W, generator=torch.Generator().manual_seed(32))
    y64, taps, _ = oracle.unet_forward_with_taps(sd, x, dtype=torch.float64)
    net = net.to(dev()).eval()
    for precision, bound in (("fp32", 1e-5), ("bf16", 2e-2)):
        net.set_precision(precision)
        y = net(x=x.to(dev()))
        plan = net.plan_for(B, H, W, dev())
        got = plan.activations()
        assert len(got) == 2 * 5 + 3 * 4, sorted(got)
        worst = max((rel_l2(t, taps[k]), k) for k, t in got.items())
        print(f"{geom} {precision}: worst tap {worst[1]} rel-L2 {worst[0]:.3e}; output rel-L2 {rel_l2(y, y64):.3e}")
        for k, t in got.items():
            assert t.shape == taps[k].shape, k
            assert rel_l2(t, taps[k]) <= bound, (precision, k, rel_l2(t, taps[k]))
        if precision == "fp32":
            err_mm = float((y.cpu().double() - y64).abs().max()) * MM_PER_UNIT
            assert float(y64.abs().max()) > 0.5 and err_mm <= 1e-3, f"max-abs depth error {err_mm:.3e} mm"
        else:
            assert rel_l2(y, y64) <= 2e-2
    net.set_precision("bf16")


def test_optimizer_and_ema_whole_arena():
    """a13 / a14 over EVERY parameter (round 1 compared one tensor at rtol 5e-2): the fused Adam + EMA kernel is fed
    the GPU path's own gradients, so the check is decoupled from bf16 gradient noise -- after each of 4 steps the
    parameter arena, both Adam moments and the EMA shadow must equal oracle.TrainOracle.apply_update (torch.optim.Adam
    with coupled L2 and the torch_ema 0.3 formula, pinned to reference-generated Adam steps in tests/test_oracle_golden.py)
    run on those same gradients."""
    from gelslim_depth_b200.train.engine import FusedTrainer
    net, sd = make_net(3, 1, 41, dims=(64, 128, 256))
    names = [k for k, _ in net.named_parameters()]
    g = torch.Generator().manual_seed(42)
    x = torch.rand(2, 3, 40, 53, generator=g).to(dev())
    tgt = (-0.9 * torch.rand(2, 1, 40, 53, generator=g)).to(dev())
    net = net.to(dev()).train()
    ft = FusedTrainer(net)
    tr = oracle.TrainOracle(sd, dtype=torch.float64)

    def arena(d):
        return torch.cat([d[k].flatten() for k in names])

    for step in range(4):
        ft.step(x, tgt)
        grads, off = {}, 0
        flat_g = ft.flat_g.double().cpu()
        for k, (o, n) in zip(names, ft.views):
            grads[k] = flat_g[o:o + n].view(sd[k].shape)
        tr.apply_update(grads)
        n = arena(tr.sd).numel()
        for what, got, want, tol in (("params", ft.flat_p, arena(tr.sd), 2e-6), ("adam m", ft.m, arena(tr.m), 1e-6),
                                     ("adam v", ft.v, arena(tr.v), 1e-6), ("ema shadow", ft.shadow, arena(tr.shadow), 2e-6)):
            err = float((got[:n].double().cpu() - want).abs().max() / (want.abs().max() + 1e-30))
            assert err <= tol, (step, what, err)
    assert ft.counter.tolist() == [4, 4] and tr.t == 4


def test_every_parameter_gradient_vs_bf16_restatement():
    """a12 with stated bounds for EVERY parameter (round 1 asserted them for the head only).  Random train-mode-BatchNorm
    nets amplify rounding noise in backward (BatchNorm backward subtracts means: heavy cancellation), so bf16 storage of
    the activation gradients puts ~15 % of zero-mean noise on early-layer gradients -- of the GPU path and of
    oracle/bf16_sim.py (the fp32 restatement with bf16 rounding at the same storage points) alike.  What an
    implementation error would look like instead is BIAS: a wrong factor, a missing term, a shifted tap.  Per parameter:
      * direction: cos(g_gpu, g_sim) >= 0.97 (measured >= 0.982);
      * scale: the projection <g_gpu, g_sim> / |g_sim|^2 lies in [0.92, 1.05] (no systematic gain error; zero-mean noise
        on g_sim itself shrinks it by |noise|^2 / |g|^2, measured 0.944 .. 1.005);
      * size of the noise: rel-L2 <= 0.3, and not larger than 1.5 x the restatement's own distance to fp32 (+ 2e-2);
    and where the backward path is short (OutConv, last BatchNorm) the gradient matches the fp32 reference to 1e-2."""
    net, sd = make_net(6, 2, 51, dims=(64, 128, 256))
    g = torch.Generator().manual_seed(52)
    x = torch.rand(2, 6, 48, 59, generator=g)
    tgt = -0.9 * torch.rand(2, 2, 48, 59, generator=g)
    _, grads_ref, _, _ = oracle.TrainOracle(sd).loss_and_grads(x, tgt)
    _, grads_q, _ = oracle.loss_and_grads_bf16(sd, x, tgt)
    net = net.to(dev()).train()
    y = net(x=x.to(dev()))
    torch.mean((y - tgt.to(dev())) ** 2).backward()
    rows = []
    for name, p in net.named_parameters():
        gg, gq, gr = p.grad.double().cpu().flatten(), grads_q[name].double().flatten(), grads_ref[name].double().flatten()
        cos = float(torch.dot(gg, gq) / (gg.norm() * gq.norm() + 1e-300))
        gain = float(torch.dot(gg, gq) / (gq.norm() ** 2 + 1e-300))
        rows.append((name, cos, gain, rel_l2(gg, gq), rel_l2(gq, gr), rel_l2(gg, gr)))
    print("min cos %.4f  gain [%.3f, %.3f]  worst rel-L2 vs sim %.3f  (sim vs fp32 up to %.3f)" % (
        min(r[1] for r in rows), min(r[2] for r in rows), max(r[2] for r in rows), max(r[3] for r in rows), max(r[4] for r in rows)))
    for name, cos, gain, e_q, e_sim, e_ref in rows:
        assert cos >= 0.97, (name, cos)
        assert 0.92 <= gain <= 1.05, (name, gain)
        assert e_q <= 0.3 and e_q <= 1.5 * e_sim + 2e-2, (name, e_q, e_sim)
        if name.startswith("outc") or name.startswith("up.1.conv.double_conv.4"):
            assert e_ref <= 1e-2, (name, e_ref)


def test_loss_curve_200_steps_full_geometry():
    """north star: matching training-loss curves over 200 steps, at the headline geometry (UNet(6,2), 6x320x427, batch 8,
    trainer init N(0, 0.01), Adam(1e-3, wd 1e-6)).  Reference curve: the fp32 restatement of the reference step
    (oracle.unet_forward in train mode + torch autograd + torch.optim.Adam, TF32 off) run with torch on this GPU --
    test infrastructure may use cuDNN, the product may not.  Both start from the same weights and see the same batches."""
    from gelslim_depth_b200.models.unet import UNet
    from gelslim_depth_b200.train.engine import FusedTrainer
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B, H, W, steps = 8, 320, 427, 200
    sd0 = oracle.trainer_init_state_dict(oracle.random_init_state_dict(6, 2, seed=7), seed=8)
    g = torch.Generator().manual_seed(9)
    xs = torch.rand(4, B, 6, H, W, generator=g).to(dev())                      # 4 batches, cycled (an "epoch" of 4)
    # a learnable target in the range of 'min_max_to_0_-1' depth maps: per finger, a 5x5 box blur of the finger's mean
    # intensity (pure noise targets would only teach the network their mean within the first steps)
    fingers = torch.stack([xs[:, :, 0:3].mean(2), xs[:, :, 3:6].mean(2)], 2)                  # (4, B, 2, H, W)
    ts = -0.9 * torch.nn.functional.avg_pool2d(fingers.flatten(0, 1), 5, stride=1, padding=2).view(4, B, 2, H, W)
    ts = ((ts + 0.45) * 6.0 - 0.45).clamp(-0.9, 0.0).contiguous()                             # stretch the contrast
    # reference
    keys = [k for k in sd0 if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    sd = {k: v.clone().to(dev()) for k, v in sd0.items()}
    leaves = [sd[k].requires_grad_(True) for k in keys]
    opt = torch.optim.Adam(leaves, lr=1e-3, weight_decay=1e-6)
    ref = []
    for i in range(steps):
        opt.zero_grad(set_to_none=True)
        out, _, stats = oracle.unet_forward_with_taps(sd, xs[i % 4], training=True, want_taps=False)
        loss = torch.mean((out - ts[i % 4]) ** 2)
        loss.backward()
        opt.step()
        with torch.no_grad():
            for prefix, (mean, var_unb) in stats.items():
                sd[prefix + ".running_mean"].mul_(0.9).add_(0.1 * mean)
                sd[prefix + ".running_var"].mul_(0.9).add_(0.1 * var_unb)
        ref.append(float(loss))
    del opt, leaves, sd
    torch.cuda.empty_cache()
    # product
    net = UNet(6, 2)
    net.load_state_dict(sd0)
    net = net.to(dev()).train()
    ft = FusedTrainer(net, use_graph=True)
    got = [float(ft.step(xs[i % 4], ts[i % 4])) for i in range(steps)]
    ft.close()
    ref_t, got_t = torch.tensor(ref).log10(), torch.tensor(got).log10()
    k = torch.ones(1, 1, 9) / 9
    sm = lambda v: torch.nn.functional.conv1d(v[None, None], k)[0, 0]
    dev_max, dev_mean = float((sm(ref_t) - sm(got_t)).abs().max()), float((sm(ref_t) - sm(got_t)).abs().mean())
    final = float(torch.tensor(got[-20:]).median() / torch.tensor(ref[-20:]).median())
    print(f"full-geometry curve: ref {ref[0]:.4f} -> {ref[-1]:.2e}, gpu {got[0]:.4f} -> {got[-1]:.2e}; smoothed log10 deviation max "
          f"{dev_max:.3f} mean {dev_mean:.3f}; final level ratio {final:.3f}")
    assert abs(got[0] - ref[0]) <= 2e-2 * ref[0]
    assert got[-1] < 0.2 * got[0] and all(map(lambda v: v == v, got))        # it trains, no NaN
    assert dev_max <= 0.2 and dev_mean <= 0.05, (dev_max, dev_mean)           # decades, after a 9-step moving average (measured 0.054 / 0.009)
    assert 0.6 <= final <= 1.6, final                                         # measured 0.88
