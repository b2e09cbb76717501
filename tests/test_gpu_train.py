"""GPU parity of the training step (train_utils/train_unet.py:346-377) against the CPU oracle / reference goldens.

Stated bf16 tolerances.  Every backward operator is checked in isolation against torch autograd on identical
bf16-rounded inputs (test_backward_ops_vs_autograd): fp32 outputs (weight / bias / gamma / beta gradients) agree to
1e-5 relative, bf16 outputs to one rounding (3e-3 relative L2).  End to end, activations and activation gradients are
stored in bf16 between kernels; with train-mode BatchNorm each layer removes the (large) DC component of a random
conv's output, which amplifies the relative bf16 noise by ~1.4x per unit on synthetic weights, so:
  * 2- and 3-level nets (7 / 13 GEMM layers): forward rel-L2 <= 2e-2, every parameter gradient rel-L2 <= 8e-2;
  * the full 5-level net (23 layers): forward rel-L2 <= 1e-1, loss within 6e-2, gradients of the last block <= 3e-2;
  * BatchNorm running statistics (momentum 0.1, unbiased variance) within 3e-2."""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def make_net(cin, ncls, seed, init="conditioned", dims=(64, 128, 256, 512, 1024)):
    from gelslim_depth_b200.models.unet import UNet
    torch.manual_seed(seed)
    net = UNet(cin, ncls, layer_dimensions=list(dims))
    fn = oracle.conditioned_state_dict if init == "conditioned" else oracle.trainer_init_state_dict
    sd = fn(net.state_dict(), seed=seed + 1)
    net.load_state_dict(sd)
    return net, sd


@pytest.mark.parametrize("cin,ncls,h,w,dims", [(3, 1, 40, 53, (64, 128)), (6, 2, 48, 59, (64, 128, 256)),
                                               (3, 1, 64, 85, (64, 128, 256, 512, 1024))])
def test_train_forward_backward_vs_oracle(cin, ncls, h, w, dims):
    full = len(dims) == 5
    net, sd = make_net(cin, ncls, 3, dims=dims)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, cin, h, w, generator=g)
    tgt = -0.9 * torch.rand(2, ncls, h, w, generator=g)
    tr = oracle.TrainOracle(sd)
    loss_ref, grads_ref, stats_ref, y_ref = tr.loss_and_grads(x, tgt)          # fp32 reference arithmetic
    loss_q, grads_q, y_q = oracle.loss_and_grads_bf16(sd, x, tgt)              # same, with the bf16 storage points
    net = net.to(dev()).train()
    y = net(x=x.to(dev()))
    assert y.requires_grad and y.shape == y_ref.shape
    fwd, fwd_q = rel_l2(y.detach(), y_ref), rel_l2(y.detach(), y_q)
    assert fwd < (1e-1 if full else 4e-2), fwd                                  # inherent bf16 distance to fp32
    loss = torch.mean((y - tgt.to(dev())) ** 2)          # the reference's MSE_loss (train_unet.py:51-52), torch autograd glue
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_ref)) < 6e-2 * abs(float(loss_ref)) + 1e-4
    errs, errs_q = {}, {}
    for name, p in net.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, name
        errs[name] = rel_l2(p.grad, grads_ref[name])
        errs_q[name] = rel_l2(p.grad, grads_q[name])
    worst, worst_q = max(errs.values()), max(errs_q.values())

    def flat(gd):
        return torch.cat([gd[k].flatten().double().cpu() for k in grads_ref])
    g_gpu = flat({k: p.grad for k, p in net.named_parameters()})
    cos_ref = float(torch.nn.functional.cosine_similarity(g_gpu, flat(grads_ref), dim=0))
    cos_q = float(torch.nn.functional.cosine_similarity(g_gpu, flat(grads_q), dim=0))
    cos_sim_ref = float(torch.nn.functional.cosine_similarity(flat(grads_q), flat(grads_ref), dim=0))
    print(f"dims={dims} fwd vs fp32 {fwd:.4f} vs bf16-sim {fwd_q:.5f}; worst grad vs fp32 {worst:.4f} vs bf16-sim {worst_q:.4f}; "
          f"cos(gpu,fp32) {cos_ref:.4f} cos(gpu,sim) {cos_q:.4f} cos(sim,fp32) {cos_sim_ref:.4f}")
    # Random train-mode-BatchNorm nets are chaotic: rounding only the WEIGHTS to bf16 moves early-layer gradients by
    # 20-80 % (oracle/bf16_sim.py).  So: (1) the GPU path is closer to the bf16 restatement than that restatement is to
    # fp32; (2) the whole-gradient direction agrees with fp32 as well as the restatement's does; (3) where the backward
    # path is short (head, last BatchNorm) the gradient matches fp32 pointwise.
    assert fwd_q < fwd and worst_q < worst, (fwd_q, fwd, worst_q, worst)
    assert cos_ref > cos_sim_ref - 0.05 and cos_ref > (0.6 if full else 0.9), (cos_ref, cos_sim_ref)
    assert abs(float(loss.detach()) - float(loss_q)) < 2e-3 * abs(float(loss_q)) + 1e-5
    last = [k for k in errs if k.startswith("outc")]
    assert max(errs[k] for k in last) < 1e-2, {k: errs[k] for k in last}      # vs fp32 where the path is short
    assert all(torch.isfinite(p.grad).all() for p in net.parameters())
    # running statistics after one train-mode forward (momentum 0.1, unbiased variance)
    got = dict(net.named_buffers())
    for prefix, (mean, var_unb) in stats_ref.items():
        rm = 0.9 * sd[prefix + ".running_mean"] + 0.1 * mean
        rv = 0.9 * sd[prefix + ".running_var"] + 0.1 * var_unb
        tol = 6e-2 if full else 3e-2
        assert torch.allclose(got[prefix + ".running_mean"].cpu(), rm, rtol=tol, atol=tol), prefix
        assert torch.allclose(got[prefix + ".running_var"].cpu(), rv, rtol=tol, atol=1e-3), prefix
        assert int(got[prefix + ".num_batches_tracked"]) == 1


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev())


def _nchw(t):
    return t.permute(0, 3, 1, 2).float().cpu()


def _bf(t):
    return t.to(torch.bfloat16).float()


def test_backward_ops_vs_autograd():
    """Each backward operator alone against torch autograd on identical bf16-rounded inputs."""
    import torch.nn.functional as F
    from gelslim_depth_b200.train import ops
    g = torch.Generator().manual_seed(7)
    d = dev()
    # ---- BatchNorm(train) + ReLU forward apply and backward
    C = 128
    z = _bf(torch.randn(2, C, 13, 17, generator=g) + 0.7)
    zt = z.clone().requires_grad_(True)
    bn = torch.nn.BatchNorm2d(C).train()
    bn.weight.data = 0.5 + torch.rand(C, generator=g)
    bn.bias.data = 0.1 * torch.randn(C, generator=g)
    a_ref = torch.relu(bn(zt))
    da = _bf(torch.randn(2, C, 13, 17, generator=g))
    a_ref.backward(da)
    mean, var = z.mean(dim=(0, 2, 3)), z.var(dim=(0, 2, 3), unbiased=False)
    rstd = torch.rsqrt(var + 1e-5)
    sc_d, sh_d = (bn.weight.data * rstd).to(d), (bn.bias.data - mean * bn.weight.data * rstd).to(d)
    a_d, _ = ops.bn_relu_apply(_nhwc(z), sc_d, sh_d)
    assert rel_l2(_nchw(a_d), a_ref.detach()) < 3e-3
    dz, sums = ops.bn_bwd(_nhwc(da), sc_d, sh_d, _nhwc(z), mean.to(d), rstd.to(d), bn.weight.data.to(d), 2 * 13 * 17)
    assert rel_l2(_nchw(dz), zt.grad) < 4e-3
    assert rel_l2(sums[C:], bn.weight.grad) < 1e-3 and rel_l2(sums[:C], bn.bias.grad) < 1e-3
    # ---- conv input gradient through the flipped-tap operand
    cin, cout = 128, 64
    xw = _bf(torch.randn(2, cin, 13, 17, generator=g)).requires_grad_(True)
    wt = _bf(torch.randn(cout, cin, 3, 3, generator=g) * 0.05)
    zz = F.conv2d(xw, wt, padding=1)
    dzz = _bf(torch.randn(zz.shape, generator=g))
    zz.backward(dzz)
    dx = ops.conv(_nhwc(dzz), ops.pack_weight(1, wt.to(d), cout, cin), cin, 9)
    assert rel_l2(_nchw(dx), xw.grad) < 3e-3
    # ---- max-pool backward (+ skip gradient), PyTorch tie rule = first maximum
    ap = _bf(torch.randint(0, 4, (2, 64, 12, 15), generator=g).float()).requires_grad_(True)     # many ties on purpose
    pp = F.max_pool2d(ap, 2)
    dp, dsk = _bf(torch.randn(pp.shape, generator=g)), _bf(torch.randn(2, 64, 12, 15, generator=g))
    pp.backward(dp)
    df = ops.maxpool_bwd(_nhwc(ap.detach()), _nhwc(dp), _nhwc(dsk))
    assert torch.equal(_nchw(df), _bf(ap.grad + dsk))
    # ---- transposed conv forward, input gradient (5-D TMA view) and weight gradient, with an F.pad row/column
    ci, co, hs, ws = 256, 128, 6, 7
    xi = _bf(torch.randn(2, ci, hs, ws, generator=g)).requires_grad_(True)
    wtt = _bf(torch.randn(ci, co, 2, 2, generator=g) * 0.05).requires_grad_(True)
    bt = 0.1 * torch.randn(co, generator=g)
    uu = F.conv_transpose2d(xi, wtt, bt, stride=2)
    du_full = _bf(torch.randn(2, co, 2 * hs + 1, 2 * ws + 1, generator=g))
    uu.backward(du_full[:, :, :2 * hs, :2 * ws])
    dufd = _nhwc(du_full)
    din = ops.convt_dgrad(dufd, (0, 0), ops.pack_weight(3, wtt.detach().to(d), co, ci), ci, hs, ws)
    assert rel_l2(_nchw(din), xi.grad) < 3e-3
    gw = torch.empty(ci, co, 2, 2, device=d)
    ops.convt_wgrad(_nhwc(xi.detach()), dufd, (0, 0), gw)
    assert rel_l2(gw, wtt.grad) < 1e-4
    uf = ops.conv(_nhwc(xi.detach()), ops.pack_weight(2, wtt.detach().to(d), co, ci), co, ntaps=1, groups=4,
                  scale=ops.ones(d, 4 * co), shift=bt.repeat(4).to(d))
    assert rel_l2(_nchw(uf), uu.detach()) < 3e-3
    # ---- 1x1 head backward
    al = _bf(torch.rand(2, 64, 9, 11, generator=g)).requires_grad_(True)
    wh = torch.randn(2, 64, 1, 1, generator=g).requires_grad_(True)
    bh = torch.randn(2, generator=g).requires_grad_(True)
    yy = F.conv2d(al, wh, bh)
    dyy = torch.randn(yy.shape, generator=g)
    yy.backward(dyy)
    dwh, dbh = torch.zeros(2, 64, device=d), torch.zeros(2, device=d)
    dal = ops.head_bwd(_nhwc(al.detach()), dyy.to(d), wh.detach().reshape(2, 64).to(d), dwh, dbh)
    assert rel_l2(_nchw(dal), al.grad) < 3e-3 and rel_l2(dwh, wh.grad.reshape(2, 64)) < 1e-5 and rel_l2(dbh, bh.grad) < 1e-5
    # ---- first-layer weight gradient
    x3 = _bf(torch.rand(2, 6, 11, 13, generator=g))
    w3 = torch.zeros(64, 6, 3, 3, requires_grad=True)
    dz3 = _bf(torch.randn(2, 64, 11, 13, generator=g))
    F.conv2d(x3, w3, padding=1).backward(dz3)
    x16 = torch.zeros(2, 11, 13, 16)
    x16[..., :6] = x3.permute(0, 2, 3, 1)
    g3 = torch.empty(64, 6, 3, 3, device=d)
    ops.wgrad_first(x16.to(torch.bfloat16).to(d), _nhwc(dz3), 6, g3)
    assert rel_l2(g3, w3.grad) < 1e-5


def test_reference_style_loop_adam_losses(golden_full):
    """The unmodified loop body of train_unet.py:346-377 (torch.optim.Adam on unet.parameters()) runs on the drop-in
    module; losses follow the reference-generated curve (trainer init N(0,0.01), fixture g1_train)."""
    from gelslim_depth_b200.models.unet import UNet
    g = golden_full["g1_train"]
    torch.manual_seed(g["module_seed"])
    net = UNet(g["cin"], g["ncls"])
    sd = oracle.trainer_init_state_dict(net.state_dict(), seed=g["init_seed"])
    assert oracle.state_dict_digest(sd) == g["digest"]
    net.load_state_dict(sd)
    net = net.to(dev()).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-6)
    x, tgt = g["x"].to(dev()), g["target"].to(dev())
    losses = []
    for _ in range(4):
        opt.zero_grad()
        out = net(x=x)
        loss = torch.mean((out - tgt) ** 2)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    ref = g["adam_losses"][:4]
    assert abs(losses[0] - ref[0]) < 2e-2 * abs(ref[0]), (losses, ref)
    for a, b in zip(losses, ref):
        assert abs(a - b) < 0.15 * abs(b) + 1e-3, (losses, ref)      # Adam's sign-like first steps amplify bf16 noise
    assert losses[-1] < losses[0]


def test_fused_trainer_matches_oracle_steps():
    """FusedTrainer (fused MSE + backward + Adam(coupled L2) + EMA arena kernel) vs oracle.TrainOracle."""
    from gelslim_depth_b200.train.engine import FusedTrainer
    net, sd = make_net(3, 1, 9)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 3, 32, 43, generator=g)
    tgt = -0.9 * torch.rand(2, 1, 32, 43, generator=g)
    tr = oracle.TrainOracle(sd)
    ref_losses = [tr.step(x, tgt) for _ in range(3)]
    net = net.to(dev()).train()
    ft = FusedTrainer(net)
    losses = [float(ft.step(x.to(dev()), tgt.to(dev()))) for _ in range(3)]
    assert abs(losses[0] - ref_losses[0]) < 2e-2 * abs(ref_losses[0]) + 1e-4, (losses, ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 0.1 * abs(b) + 1e-3, (losses, ref_losses)
    # EMA shadow follows shadow -= (1-d)(shadow - p), d = min(0.995, (1+n)/(10+n)): compare one tensor with the oracle
    k = "outc.conv.bias"
    off, n = ft.views[-1]
    assert torch.allclose(ft.shadow[off:off + n].cpu(), tr.shadow[k].flatten(), rtol=5e-2, atol=2e-3)
    assert torch.allclose(dict(net.named_parameters())[k].detach().cpu().flatten(), tr.sd[k].flatten(), rtol=5e-2, atol=2e-3)


def test_loss_curve_200_steps_vs_reference():
    """north star: matching training-loss curves over 200 steps.  Fixture: 200 steps of the unmodified reference
    (fp32, CPU, torch.optim.Adam lr 1e-3 wd 1e-6, trainer init) on a fixed synthetic problem
    (tests/golden/make_train_curve.py).  The B200 path (bf16, fused MSE + Adam + EMA) must follow it: same first-step
    loss, the same decades-long descent and the same final level.  Training is chaotic and the GPU path is not bit-
    reproducible (fp32 atomics in the BatchNorm statistics and the split-K weight gradients), so the criteria are robust
    statistics; 50 repetitions (tests/stress_train_curve.py) gave: log10 distance of the 10-step-smoothed curves max
    0.04-0.39 (at the loss spike around step 105-115 that the reference has too, or at an isolated late spike), mean
    0.01-0.10, 90th percentile 0.02-0.29, median of the last 20 losses 0.83-2.05 x the reference's.  Thresholds (about
    2.5x the observed extremes): max <= 1.0 decade, mean <= 0.25, 90th percentile <= 0.6, median of the last 20 losses
    within a factor 4."""
    import math
    import os
    from gelslim_depth_b200.models.unet import UNet
    from gelslim_depth_b200.train.engine import FusedTrainer
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "train_curve.pt"), weights_only=False)
    torch.manual_seed(g["module_seed"])
    net = UNet(3, 1)
    sd = oracle.trainer_init_state_dict(net.state_dict(), seed=g["init_seed"])
    assert oracle.state_dict_digest(sd) == g["digest"]
    net.load_state_dict(sd)
    net = net.to(dev()).train()
    ft = FusedTrainer(net)
    X, T = g["X"].to(dev()), g["T"].to(dev())
    losses = []
    for step in range(g["steps"]):
        idx = torch.arange(4) + 4 * (step % 4)
        losses.append(ft.step(X[idx], T[idx]))
    losses = [float(v) for v in torch.cat(losses).cpu()]
    ref = g["losses"]
    assert abs(losses[0] - ref[0]) < 1e-2 * ref[0], (losses[0], ref[0])

    def smooth(v):
        return [sum(v[max(0, i - 9): i + 1]) / len(v[max(0, i - 9): i + 1]) for i in range(len(v))]
    a, b = smooth(losses), smooth(ref)
    dist = [abs(math.log10(x) - math.log10(y)) for x, y in zip(a, b)]
    print("gpu ", [f"{v:.2e}" for v in losses[::20]], f"{losses[-1]:.2e}")
    print("ref ", [f"{v:.2e}" for v in ref[::20]], f"{ref[-1]:.2e}")
    print("max/mean log10 distance of smoothed curves:", max(dist), sum(dist) / len(dist))
    p90 = sorted(dist)[int(0.9 * len(dist))]
    assert max(dist) <= 1.0 and sum(dist) / len(dist) <= 0.25 and p90 <= 0.6, (max(dist), sum(dist) / len(dist), p90)
    med = lambda v: sorted(v[-20:])[10]      # noqa: E731  (robust to an isolated late spike)
    assert 0.25 * med(ref) < med(losses) < 4.0 * med(ref), (med(losses), med(ref))
    assert med(losses) < 1e-2 * losses[0]


def test_ema_average_parameters_and_checkpoint_roundtrip(tmp_path):
    """eval under `ema.average_parameters()` and state_dict save/load (train_unet.py:380-390,476-484)."""
    from gelslim_depth_b200.models.unet import UNet
    from gelslim_depth_b200.train.engine import FusedTrainer
    net, sd = make_net(3, 1, 5, dims=(64, 128))
    net = net.to(dev()).train()
    ft = FusedTrainer(net)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(2, 3, 32, 43, generator=g).to(dev())
    t = (-0.9 * torch.rand(2, 1, 32, 43, generator=g)).to(dev())
    for _ in range(3):
        ft.step(x, t)
    live = ft.flat_p.clone()
    net.eval()
    y_live = net(x=x)
    with ft.average_parameters():
        assert torch.equal(ft.flat_p, ft.shadow)
        y_ema = net(x=x)                                  # packed-weight cache must notice the in-place swap
        torch.save(net.state_dict(), tmp_path / "ema.pth")
    assert torch.equal(ft.flat_p, live)
    assert not torch.allclose(y_live, y_ema)
    assert torch.allclose(net(x=x), y_live)
    net2 = UNet(3, 1, layer_dimensions=[64, 128])
    net2.load_state_dict(torch.load(tmp_path / "ema.pth", map_location="cpu"))
    net2 = net2.to(dev()).eval()
    assert torch.allclose(net2(x=x), y_ema)


def test_fused_trainer_cuda_graph_matches_eager():
    """use_graph=True replays the captured step; losses must match the eager trainer step for step."""
    from gelslim_depth_b200.train.engine import FusedTrainer
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 3, 32, 43, generator=g).to(dev())
    tgt = (-0.9 * torch.rand(2, 1, 32, 43, generator=g)).to(dev())
    curves = []
    for use_graph in (False, True):
        net, _ = make_net(3, 1, 9, dims=(64, 128, 256))
        net = net.to(dev()).train()
        ft = FusedTrainer(net, use_graph=use_graph)
        curves.append([float(ft.step(x, tgt)) for _ in range(6)])
        assert int(ft.counter[0]) == 6 and int(net.inc.double_conv[1].num_batches_tracked) == 6
    for a, b in zip(*curves):
        assert abs(a - b) < 8e-2 * abs(b) + 1e-5, curves       # atomics order differs run to run (chaotic nets); same trajectory
    assert curves[1][-1] < curves[1][0]


def _blocks(net):
    enc = [net.inc.double_conv] + [d.maxpool_conv[1].double_conv for d in net.down]
    dec = [(u.up, u.conv.double_conv) for u in net.up]
    return enc, dec


class _PackedTrainWeights:
    """test driver of gsd_op_pack_weights_batched: the same item table csrc/train_plan.h builds (forward + dgrad operands of
    every layer, one launch), with the operands exposed per layer"""

    def __init__(self, net):
        from gelslim_depth_b200._lib import PackItem, lib
        from gelslim_depth_b200.train import ops
        enc, dec = _blocks(net)
        self.fwd, self.dgrad = {}, {}
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
        device = items[0][1].device
        al = lambda n: (n + 7) // 8 * 8                                  # every operand 16-byte aligned
        elems = sum(al(ops.pack_out_elems(m, O, I, ip)) + (al(ops.pack_out_elems(1, O, I)) if dg else 0)
                    for (m, _, O, I, ip, _, dg) in items)
        self.arena = torch.empty(elems, dtype=torch.bfloat16, device=device)
        table = (PackItem * len(items))()
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
        self.n_items, self.total = len(items), units
        self.table = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(device)
        self.repack()

    def repack(self):
        from gelslim_depth_b200.train import ops
        ops.pack_weights_batched(self.table, self.n_items, self.total, self.arena.device)


def test_batched_weight_pack_matches_per_layer_pack():
    """gsd_op_pack_weights_batched (one launch, tiled, forward + dgrad operands from one staged tile) is bit-identical
    to the per-layer gsd_op_pack_weight modes 0-3 for every layer, including the 6->16 channel padding of inc.0."""
    from gelslim_depth_b200.train import ops
    net, _ = make_net(6, 2, 11, dims=(64, 128, 256))
    net = net.to(dev()).train()
    pw = _PackedTrainWeights(net)
    enc, dec = _blocks(net)
    n = 0
    for bi, seq in enumerate(enc):
        for ci in (0, 3):
            w = seq[ci].weight.detach()
            O, I = w.shape[:2]
            first = bi == 0 and ci == 0
            assert torch.equal(pw.fwd[id(seq[ci])], ops.pack_weight(0, w, O, I, 16 if first else I))
            if not first:
                assert torch.equal(pw.dgrad[id(seq[ci])], ops.pack_weight(1, w, O, I))
            n += 1
    for up, seq in dec:
        w = up.weight.detach()
        I, O = w.shape[:2]
        assert torch.equal(pw.fwd[id(up)], ops.pack_weight(2, w, O, I))
        assert torch.equal(pw.dgrad[id(up)], ops.pack_weight(3, w, O, I))
        for ci in (0, 3):
            cw = seq[ci].weight.detach()
            assert torch.equal(pw.fwd[id(seq[ci])], ops.pack_weight(0, cw, cw.shape[0], cw.shape[1]))
            assert torch.equal(pw.dgrad[id(seq[ci])], ops.pack_weight(1, cw, cw.shape[0], cw.shape[1]))
    assert n == 6
    # in-place parameter update + repack
    with torch.no_grad():
        net.inc.double_conv[3].weight.mul_(2.0)
    pw.repack()
    w = net.inc.double_conv[3].weight.detach()
    assert torch.equal(pw.dgrad[id(net.inc.double_conv[3])], ops.pack_weight(1, w, 64, 64))


@pytest.mark.parametrize("ncls", [1, 2])
def test_fused_last_unit_matches_unfused_ops(ncls):
    """gsd_op_bn_relu_head_fwd / gsd_op_head_bn_bwd (last unit without a stored activation) against the chain of
    bn_relu_apply -> head_fwd and head_bwd -> bn_bwd on materialised tensors: identical roundings, fp32 sums differ only
    in accumulation order."""
    from gelslim_depth_b200.train import ops
    d = dev()
    g = torch.Generator().manual_seed(20 + ncls)
    B, H, W, C = 3, 21, 27, 64
    z = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16).to(d)
    scale = (torch.rand(C, generator=g) + 0.5).to(d) * torch.where(torch.arange(C) % 7 == 0, -1.0, 1.0).to(d)
    shift = (0.2 * torch.randn(C, generator=g)).to(d)
    mean = (0.1 * torch.randn(C, generator=g)).to(d)
    rstd = (torch.rand(C, generator=g) + 0.5).to(d)
    gamma = (torch.rand(C, generator=g) + 0.5).to(d)
    w = (0.2 * torch.randn(ncls, C, generator=g)).to(d)
    bias = torch.randn(ncls, generator=g).to(d)
    dy = torch.randn(B, ncls, H, W, generator=g).to(d)
    # forward
    a, _ = ops.bn_relu_apply(z, scale, shift)
    y_ref = ops.head_fwd(a, w, bias)
    y = ops.bn_relu_head_fwd(z, scale, shift, w, bias)
    assert torch.equal(y, y_ref)
    # backward
    dw_ref, db_ref = torch.zeros(ncls, C, device=d), torch.zeros(ncls, device=d)
    da = ops.head_bwd(a, dy, w, dw_ref, db_ref)
    dz_ref, sums_ref = ops.bn_bwd(da, scale, shift, z, mean, rstd, gamma, B * H * W)
    dw, db = torch.zeros(ncls, C, device=d), torch.zeros(ncls, device=d)
    dz, sums = ops.head_bn_bwd(z, dy, w, scale, shift, mean, rstd, gamma, dw, db)
    torch.cuda.synchronize()
    assert torch.allclose(dw, dw_ref, rtol=1e-5, atol=1e-4) and torch.allclose(db, db_ref, rtol=1e-5, atol=1e-4)
    assert torch.allclose(sums, sums_ref, rtol=1e-5, atol=1e-4)
    diff = (dz.float() - dz_ref.float()).abs()
    assert float((diff > 2 ** -7 * dz_ref.float().abs() + 1e-6).float().mean()) == 0.0   # at most one bf16 ulp (sums order)


def test_train_step_through_the_c_abi_only():
    """gsd_train_plan_create / bind / gsd_train_step called through ctypes alone -- torch tensors are nothing but
    the memory here (flat arenas the caller owns).  Three steps of train_unet.py:346-377 against oracle.TrainOracle, the
    gradients of step 1 against the oracle's, and gsd_train_forward / gsd_backward / gsd_adam_ema_step called separately
    against the one-call form."""
    import ctypes as C
    from gelslim_depth_b200 import _lib
    from gelslim_depth_b200._lib import lib
    cin, ncls, dims, B, H, W = 3, 1, (64, 128, 256), 2, 40, 53
    net, sd = make_net(cin, ncls, 21, dims=dims)
    names = [k for k, _ in net.named_parameters()]
    g = torch.Generator().manual_seed(2)
    x = torch.rand(B, cin, H, W, generator=g)
    tgt = -0.9 * torch.rand(B, ncls, H, W, generator=g)
    tr = oracle.TrainOracle(sd)
    _, grads_ref, _, _ = tr.loss_and_grads(x, tgt)
    ref_losses = [tr.step(x, tgt) for _ in range(3)]

    geo = _lib.Geometry()
    geo.batch, geo.in_channels, geo.height, geo.width, geo.n_classes, geo.n_dims = B, cin, H, W, ncls, len(dims)
    for i, d in enumerate(dims):
        geo.dims[i] = d
    geo.dtype, geo.mode = _lib.DTYPE_BF16, _lib.MODE_TRAIN

    def build():
        """plan + caller-owned memory, initialised from the state_dict"""
        h = C.c_void_p()
        assert lib.gsd_train_plan_create(C.byref(h), C.byref(geo), 0) == 0, lib.gsd_last_error()
        n = lib.gsd_train_plan_num_params(h)
        numel = (C.c_longlong * n)()
        assert lib.gsd_train_plan_param_numel(h, numel, n) == n == len(names)
        total = sum(numel)
        mem = {k: torch.zeros(total, device=dev()) for k in ("p", "g", "m", "v")}
        offs, off = [], 0
        for k, cnt in zip(names, numel):
            assert sd[k].numel() == cnt, k
            mem["p"][off:off + cnt].copy_(sd[k].flatten())
            offs.append(off)
            off += cnt
        mem["ema"] = mem["p"].clone()
        bn_keys = [k[:-len(".running_mean")] for k in sd if k.endswith(".running_mean")]
        bn = [sd[k + s].clone().to(dev()) for k in bn_keys for s in (".running_mean", ".running_var")]
        nbt = [torch.zeros((), dtype=torch.int64, device=dev()) for _ in bn_keys]
        assert lib.gsd_train_plan_num_bn(h) == len(bn_keys)
        ws = torch.empty(lib.gsd_train_plan_workspace_bytes(h), dtype=torch.uint8, device=dev())
        vp = lambda ts: (C.c_void_p * len(ts))(*ts)
        assert lib.gsd_train_plan_bind(h, vp([mem["p"].data_ptr() + 4 * o for o in offs]), vp([mem["g"].data_ptr() + 4 * o for o in offs]),
                                       vp([t.data_ptr() for t in bn]), vp([t.data_ptr() for t in nbt]), C.c_void_p(ws.data_ptr())) == 0, \
            lib.gsd_last_error()
        counter = torch.zeros(2, dtype=torch.int64, device=dev())
        opt = _lib.OptimizerState()
        opt.params, opt.grads, opt.m, opt.v, opt.ema = (mem[k].data_ptr() for k in ("p", "g", "m", "v", "ema"))
        opt.n, opt.counter = total, counter.data_ptr()
        opt.hp.lr, opt.hp.beta1, opt.hp.beta2, opt.hp.eps, opt.hp.weight_decay, opt.hp.ema_decay, opt.hp.grad_scale = \
            1e-3, 0.9, 0.999, 1e-8, 1e-6, 0.995, 1.0
        return h, mem, offs, numel, bn, nbt, ws, counter, opt

    xd, td = x.to(dev()), tgt.to(dev())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # ---- the one-call form
    h, mem, offs, numel, bn, nbt, ws, counter, opt = build()
    loss = torch.zeros(1, device=dev())
    losses = []
    for step in range(3):
        assert lib.gsd_train_step(h, xd.data_ptr(), td.data_ptr(), loss.data_ptr(), C.byref(opt), st, _lib.NULL_CB, None) == 0, lib.gsd_last_error()
        losses.append(float(loss))
        if step == 0:
            g1 = mem["g"].clone()
    torch.cuda.synchronize()
    assert abs(losses[0] - ref_losses[0]) < 2e-2 * abs(ref_losses[0]) + 1e-4, (losses, ref_losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 0.1 * abs(b) + 1e-3, (losses, ref_losses)
    assert counter.tolist() == [3, 3] and all(int(t) == 3 for t in nbt)
    for k, o, cnt in zip(names, offs, numel):                 # where the backward path is short the gradient matches fp32 pointwise
        if k.startswith("outc"):
            assert rel_l2(g1[o:o + cnt], grads_ref[k].flatten()) < 1e-2, k
    flat_ref = torch.cat([grads_ref[k].flatten() for k in names]).double()
    assert float(torch.nn.functional.cosine_similarity(g1.double().cpu(), flat_ref, dim=0)) > 0.9
    assert 95 <= lib.gsd_train_plan_launches(h) <= 99      # 10 conv units + 2 transposed convs (99 when both concat dgrads take two launches)
    # ---- forward / loss / backward / optimizer as separate calls: the same first step
    h2, mem2, _, _, bn2, nbt2, ws2, counter2, opt2 = build()
    y = torch.empty(B, ncls, H, W, device=dev())
    dy = torch.empty_like(y)
    loss2 = torch.zeros(1, device=dev())
    assert lib.gsd_train_forward(h2, xd.data_ptr(), y.data_ptr(), st) == 0, lib.gsd_last_error()
    assert lib.gsd_op_mse(y.data_ptr(), td.data_ptr(), y.numel(), loss2.data_ptr(), dy.data_ptr(), st) == 0
    assert lib.gsd_backward(h2, dy.data_ptr(), st, _lib.NULL_CB, None) == 0, lib.gsd_last_error()
    g2 = mem2["g"].clone()
    assert lib.gsd_adam_ema_step(opt2.params, opt2.grads, opt2.m, opt2.v, opt2.ema, opt2.n, C.byref(opt2.hp), opt2.counter, st) == 0
    torch.cuda.synchronize()
    assert abs(float(loss2) - losses[0]) < 1e-3 * abs(losses[0])          # BatchNorm statistics use fp32 atomics: not bit-reproducible
    # two runs of the SAME path: fp32 atomics reorder the BatchNorm sums and these random nets amplify that (measured 0.02 .. 0.11)
    assert rel_l2(g2, g1) < 0.25 and counter2.tolist() == [1, 1]
    y_ref = oracle.unet_forward(sd, x, training=True)
    assert rel_l2(y, y_ref) < 4e-2
    lib.gsd_train_plan_destroy(h)
    lib.gsd_train_plan_destroy(h2)
