"""GPU parity tests (run with `-m gpu` on a B200): every CUDA path is called through the C ABI and
compared with the CPU oracle / the reference-generated golden fixtures.

Stated tolerances (bf16 mode): operands are rounded to bf16, products accumulate in fp32 (TMEM) and
each layer's output is rounded to bf16 once.  Single op vs an fp32 CPU conv on the SAME bf16-rounded
operands: |err| <= 2^-7 * |y| + 1e-3 (one bf16 ulp: fp32 summation order can flip the output rounding).  Whole network (23 layers) vs the fp32
reference: per-tensor relative L2 <= 2e-2, max-abs depth error <= 5e-2 in network units on
conditioned O(1) activations."""
import types

import pytest
import torch
import torch.nn.functional as F

import oracle

pytestmark = pytest.mark.gpu

TAPS3 = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]


def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return torch.device("cuda:0")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def bf16r(t):
    return t.to(torch.bfloat16).float()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def pack_w3(w, cin_pad=None):
    """(O,I,3,3) -> bf16 [O][9][Ipad] == csrc/elementwise.cuh pack_conv_weight_kernel"""
    O, I = w.shape[:2]
    cin_pad = cin_pad or I
    out = torch.zeros(O, 9, cin_pad)
    out[:, :, :I] = w.reshape(O, I, 9).permute(0, 2, 1)
    return out.reshape(O, 9 * cin_pad).to(torch.bfloat16)


def check_close(got, ref, what):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    tol = ref.abs() * 2 ** -7 + 1e-3
    bad = int((err > tol).sum())
    assert bad == 0, f"{what}: {bad}/{err.numel()} beyond tolerance, max err {float(err.max()):.4g}, rel_l2 {rel_l2(got, ref):.3g}"


@pytest.mark.parametrize("cin,cout,h,w,b,bn", [(64, 64, 19, 23, 2, 0), (64, 128, 16, 32, 1, 128), (128, 64, 9, 150, 1, 64),
                                               (256, 256, 10, 13, 3, 256), (128, 256, 20, 26, 1, 64), (64, 64, 40, 53, 2, 64)])
def test_conv3x3_op(cin, cout, h, w, b, bn):
    from gelslim_depth_b200.engine import conv_op
    g = torch.Generator().manual_seed(cin * 7 + cout + h)
    x = bf16r(torch.randn(b, cin, h, w, generator=g))
    wt = bf16r(torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5)
    scale = 0.5 + torch.rand(cout, generator=g)
    shift = torch.randn(cout, generator=g) * 0.3
    ref = torch.relu(F.conv2d(x, wt, padding=1) * scale[None, :, None, None] + shift[None, :, None, None])
    d = dev()
    out = conv_op(nhwc(x).to(torch.bfloat16).to(d), pack_w3(wt).to(d), scale.to(d), shift.to(d), TAPS3, relu=True, block_n=bn)
    torch.cuda.synchronize()
    check_close(out.permute(0, 3, 1, 2), ref, f"conv3x3 {cin}->{cout} {h}x{w}")


def test_first_layer_op_padded_channels():
    from gelslim_depth_b200.engine import conv_op
    g = torch.Generator().manual_seed(5)
    for cin in (3, 6):
        x = bf16r(torch.rand(2, cin, 21, 37, generator=g))
        wt = bf16r(torch.randn(64, cin, 3, 3, generator=g) * 0.3)
        ref = torch.relu(F.conv2d(x, wt, padding=1))
        x16 = torch.zeros(2, 21, 37, 16)
        x16[..., :cin] = nhwc(x)
        d = dev()
        out = conv_op(x16.to(torch.bfloat16).to(d), pack_w3(wt, 16).to(d), torch.ones(64, device=d),
                      torch.zeros(64, device=d), TAPS3, relu=True)
        torch.cuda.synchronize()
        check_close(out.permute(0, 3, 1, 2), ref, f"first layer cin={cin}")


@pytest.mark.parametrize("h,w", [(16, 32), (21, 27), (40, 53)])
def test_conv_with_fused_maxpool(h, w):
    from gelslim_depth_b200.engine import conv_op
    g = torch.Generator().manual_seed(h + w)
    x = bf16r(torch.randn(2, 64, h, w, generator=g))
    wt = bf16r(torch.randn(128, 64, 3, 3, generator=g) * (2.0 / 576) ** 0.5)
    ref = bf16r(torch.relu(F.conv2d(x, wt, padding=1)))
    d = dev()
    out, pooled = conv_op(nhwc(x).to(torch.bfloat16).to(d), pack_w3(wt).to(d), torch.ones(128, device=d),
                          torch.zeros(128, device=d), TAPS3, relu=True, pool=True)
    torch.cuda.synchronize()
    check_close(out.permute(0, 3, 1, 2), ref, "conv before pool")
    # the pooled tensor must be EXACTLY max_pool2d (floor mode, unet.py:26) of the stored bf16 tensor
    want = F.max_pool2d(out.permute(0, 3, 1, 2).float().cpu(), 2)
    assert torch.equal(pooled.permute(0, 3, 1, 2).float().cpu(), want)


@pytest.mark.parametrize("hs,ws,h,w", [(5, 6, 10, 13), (10, 13, 20, 26), (8, 8, 17, 17)])
def test_virtual_concat_with_pad(hs, ws, h, w):
    """conv over torch.cat([skip, F.pad(up)], 1) without materialising it (unet.py:43-48)."""
    from gelslim_depth_b200.engine import conv_op
    g = torch.Generator().manual_seed(hs * 31 + w)
    skip = bf16r(torch.randn(2, 64, h, w, generator=g))
    up = bf16r(torch.randn(2, 64, 2 * hs, 2 * ws, generator=g))
    dy, dx = h - 2 * hs, w - 2 * ws
    cat = torch.cat([skip, F.pad(up, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])], dim=1)
    wt = bf16r(torch.randn(64, 128, 3, 3, generator=g) * (2.0 / 1152) ** 0.5)
    ref = torch.relu(F.conv2d(cat, wt, padding=1))
    d = dev()
    out = conv_op(nhwc(skip).to(torch.bfloat16).to(d), pack_w3(wt).to(d), torch.ones(64, device=d), torch.zeros(64, device=d),
                  TAPS3, relu=True, src1=nhwc(up).to(torch.bfloat16).to(d), off=(dy // 2, dx // 2))
    torch.cuda.synchronize()
    check_close(out.permute(0, 3, 1, 2), ref, "virtual concat")


@pytest.mark.parametrize("cin,h,w,bn", [(128, 10, 13, 64), (256, 20, 26, 128), (512, 5, 6, 0)])
def test_transposed_conv_scatter(cin, h, w, bn):
    """ConvTranspose2d(k=2, s=2) + bias as a GEMM with a strided 4-view TMA scatter (unet.py:36)."""
    from gelslim_depth_b200.engine import conv_op
    cout = cin // 2
    g = torch.Generator().manual_seed(cin + h)
    x = bf16r(torch.randn(2, cin, h, w, generator=g))
    wt = bf16r(torch.randn(cin, cout, 2, 2, generator=g) * (1.0 / cin) ** 0.5)
    bias = torch.randn(cout, generator=g) * 0.2
    ref = F.conv_transpose2d(x, wt, bias, stride=2)
    packed = wt.permute(2, 3, 1, 0).reshape(4 * cout, cin).to(torch.bfloat16)     # [(dy*2+dx)*O + o][I]
    d = dev()
    out = conv_op(nhwc(x).to(torch.bfloat16).to(d), packed.to(d), torch.ones(4 * cout, device=d), bias.repeat(4).to(d),
                  [(0, 0)], relu=False, groups=4, cout=cout, block_n=bn)
    torch.cuda.synchronize()
    check_close(out.permute(0, 3, 1, 2), ref, "transposed conv")


def build_net(g, device):
    from gelslim_depth_b200.models.unet import UNet
    torch.manual_seed(g["module_seed"])
    net = UNet(g["cin"], g["ncls"], layer_dimensions=g["dims"])
    init = oracle.conditioned_state_dict if g["init"] == "conditioned" else oracle.trainer_init_state_dict
    sd = init(net.state_dict(), seed=g["init_seed"])
    assert oracle.state_dict_digest(sd) == g["digest"]
    net.load_state_dict(sd)
    return net.to(device).eval(), sd


@pytest.mark.parametrize("tag", ["g2_eval", "g3_eval"])
def test_unet_forward_vs_reference_golden(golden_full, tag):
    g = golden_full[tag]
    net, sd = build_net(g, dev())
    y = net(x=g["x"].to(dev()))
    torch.cuda.synchronize()
    assert y.shape == g["y"].shape and y.dtype == torch.float32
    ref = g["y"]
    assert float(ref.std()) > 0.05, "fixture must not be degenerate"
    assert rel_l2(y, ref) < 2e-2, rel_l2(y, ref)
    assert float((y.cpu() - ref).abs().max()) < 5e-2 * max(1.0, float(ref.abs().max()))


def test_unet_forward_vs_oracle_odd_geometry():
    """Geometry that exercises floor pooling and the right/bottom zero pad at every level (427-like)."""
    from gelslim_depth_b200.models.unet import UNet
    torch.manual_seed(1)
    net = UNet(6, 2)
    sd = oracle.conditioned_state_dict(net.state_dict(), seed=3)
    net.load_state_dict(sd)
    net = net.to(dev()).eval()
    x = torch.rand(3, 6, 53, 75, generator=torch.Generator().manual_seed(2))
    ref = oracle.unet_forward(sd, x)
    y = net(x=x.to(dev()))
    assert rel_l2(y, ref) < 2e-2
    # weight-cache invalidation: in-place parameter update (what Adam / ema.average_parameters() do)
    with torch.no_grad():
        net.outc.conv.bias.add_(1.0)
    y2 = net(x=x.to(dev()))
    assert torch.allclose(y2, y + 1.0, atol=1e-5)
    # ... and the writes torch's version counters do NOT see: torch_ema 0.3's copy_to / restore and the reference's weight
    # init go through `param.data` (train_unet.py:248-250,389,428,480); the packed-operand cache is keyed on a device-side
    # content fingerprint, so the next eval forward still runs the new weights, and an unchanged model does not re-pack
    assert int(net._pack_state[1]) == 1                      # the add_ above was seen as a change ...
    y2b = net(x=x.to(dev()))
    assert int(net._pack_state[1]) == 0 and torch.equal(y2b, y2)      # ... and nothing changed since
    v0 = net.outc.conv.bias._version
    net.outc.conv.bias.data.copy_(net.outc.conv.bias.data - 1.0)
    assert net.outc.conv.bias._version == v0                 # invisible to the host
    y3 = net(x=x.to(dev()))
    assert int(net._pack_state[1]) == 1 and torch.allclose(y3, y, atol=1e-5)
    bn = net.inc.double_conv[1]
    bn.running_mean.data.add_(0.25)                          # raw write to a BatchNorm buffer (what the training kernels do)
    sd2 = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    y4 = net(x=x.to(dev()))
    assert rel_l2(y4, oracle.unet_forward(sd2, x)) < 2e-2 and not torch.allclose(y4, y, atol=1e-3)
    # serving mode: the caller promises frozen weights, the check is skipped (documented contract)
    net.frozen_weights(True)
    bn.running_mean.data.sub_(0.25)
    assert torch.equal(net(x=x.to(dev())), y4)
    net.frozen_weights(False)
    assert torch.allclose(net(x=x.to(dev())), y, atol=1e-5)


def shipped_cfg(size, tactile_spelling=False):
    c = types.SimpleNamespace(input_tactile_image_size=size, interp_method="area", norm_scale=0.9,
                              depth_normalization_method="min_max_to_0_-1",
                              depth_normalization_parameters=(-1.9180814027786255, 0.0))
    pre = "tactile" if tactile_spelling else "image"
    setattr(c, pre + "_normalization_method", "0_255_to_0_1")
    setattr(c, pre + "_normalization_parameters", None)
    return c


@pytest.mark.parametrize("spelling", [False, True])
def test_predict_depth_pipeline_g3(spelling):
    """predict_depth_from_RGB with area down/up-sampling (G3 pipeline, complete_prediction.py:4-10)."""
    from gelslim_depth_b200.models.unet import UNet
    from gelslim_depth_b200.processing_utils.complete_prediction import predict_depth_from_RGB, predict_depth_from_frames
    torch.manual_seed(4)
    net = UNet(3, 1)
    sd = oracle.conditioned_state_dict(net.state_dict(), seed=8)
    net.load_state_dict(sd)
    net = net.to(dev()).eval()
    g = torch.Generator().manual_seed(6)
    raw = torch.randint(0, 256, (2, 6, 64, 85), generator=g).float()
    base = torch.randint(0, 256, (1, 6, 64, 85), generator=g).float()
    cfg = shipped_cfg((32, 43), spelling)
    fingers = oracle.split_fingers(oracle.get_difference_image(raw, base))
    ref = oracle.predict_depth_from_RGB(fingers, lambda t: oracle.unet_forward(sd, t), (64, 85), cfg)
    got = predict_depth_from_RGB(fingers.to(dev()), net, (64, 85), cfg)
    assert got.shape == ref.shape
    scale_mm = 1.9180814027786255 / 0.9
    assert float((got.cpu() - ref).abs().max()) < 5e-2 * scale_mm * max(1.0, float((ref / scale_mm).abs().max()))
    assert rel_l2(got, ref) < 2e-2
    # raw frames + base: difference image fused too (Left finger = channels 0:3)
    got2 = predict_depth_from_frames(raw[:, 0:3].contiguous().to(dev()), base[:, 0:3].contiguous().to(dev()), net, (64, 85), cfg)
    assert rel_l2(got2, ref[:2]) < 2e-2


def test_processing_helpers_vs_golden(golden_processing):
    from gelslim_depth_b200.processing_utils import image_utils, normalization_utils
    p = golden_processing
    d = dev()
    diff = image_utils.get_difference_image(p["raw"].to(d), p["base"].to(d))
    assert torch.equal(diff.cpu(), p["diff"])
    g = torch.Generator().manual_seed(9)
    torch.randint(0, 256, (2, 6, 64, 85), generator=g)
    torch.randint(0, 256, (1, 6, 64, 85), generator=g)
    img = torch.rand(1, 1, 320, 427, generator=g)
    down = image_utils.sample_multi_channel_image_to_desired_size(img.to(d), (160, 213), "area")
    assert torch.allclose(down.cpu(), p["area_down"], rtol=1e-5, atol=1e-6)
    p4 = ([1.0, 2.0, 3.0], [200.0, 210.0, 220.0], [100.0, 110.0, 120.0], [50.0, 60.0, 70.0])
    for m in ("mean_std", "0_255_to_-1_1", "0_255_to_0_1"):
        got = normalization_utils.normalize_tactile_image(p["norm_in"].to(d), m, 0.9, p4)
        assert torch.allclose(got.cpu(), p["norm_img_" + m], rtol=1e-5, atol=1e-5), m
    dp = (-1.9180814027786255, 0.0, -0.4, 0.3)
    for m in ("min_max_to_-1_1", "mean_std", "min_max_to_0_1", "min_max_to_0_-1"):
        got = normalization_utils.denormalize_depth_image(p["depth_in"].to(d), m, 0.9, dp)
        assert torch.allclose(got.cpu(), p["denorm_depth_" + m], rtol=1e-5, atol=1e-5), m
        got = normalization_utils.normalize_depth_image(p["depth_in"].to(d), m, 0.9, dp)
        assert torch.allclose(got.cpu(), p["norm_depth_" + m], rtol=1e-5, atol=1e-5), m


def test_forward_host_pipelined_matches_device_path():
    from gelslim_depth_b200.models.unet import UNet
    from gelslim_depth_b200.engine import make_prepost
    torch.manual_seed(2)
    net = UNet(6, 2)
    net.load_state_dict(oracle.conditioned_state_dict(net.state_dict(), seed=1))
    net = net.to(dev()).eval()
    x = torch.rand(5, 6, 32, 43, generator=torch.Generator().manual_seed(3))
    y_ref = net(x=x.to(dev())).cpu()
    plan = net.plan_for(5, 32, 43, dev())
    plan.set_chunk(2)
    pp = make_prepost(6, (32, 43), (32, 43))
    xh, yh = x.pin_memory(), torch.empty(5, 2, 32, 43).pin_memory()
    xd, yd = torch.empty_like(x, device=dev()), torch.empty(5, 2, 32, 43, device=dev())
    plan.forward_host(xh, None, pp, yh, xd, yd, net.packed_weights(plan))
    assert torch.equal(yh, y_ref)
    one_chunk_launches = plan.launches
    # ramp: smaller first / last chunk (1 + 2 + 1 + 1 frames) must cover every frame exactly once
    yh.zero_()
    plan.set_chunk(2, first=1, last=1)
    assert plan.launches == 4 * (one_chunk_launches // 3)       # 3 chunks before, 4 now
    plan.forward_host(xh, None, pp, yh, xd, yd, net.packed_weights(plan))
    assert torch.equal(yh, y_ref)


def test_forward_host_async_rotating_slots_match_device_path():
    """gsd_forward_host_async: six different batches through two rotating staging slots (each slot re-used three
    times without a host wait in between for the first four) land the same bytes as the blocking device path."""
    from gelslim_depth_b200.models.unet import UNet
    from gelslim_depth_b200.engine import make_prepost
    torch.manual_seed(2)
    net = UNet(6, 2)
    net.load_state_dict(oracle.conditioned_state_dict(net.state_dict(), seed=1))
    net = net.to(dev()).eval()
    nb, B = 6, 5
    xs = [torch.rand(B, 6, 32, 43, generator=torch.Generator().manual_seed(10 + k)) for k in range(nb)]
    refs = [net(x=x.to(dev())).cpu() for x in xs]
    assert not torch.equal(refs[0], refs[1])
    plan = net.plan_for(B, 32, 43, dev())
    packed = net.packed_weights(plan)
    pp = make_prepost(6, (32, 43), (32, 43))
    xh = [x.pin_memory() for x in xs]
    yh = [torch.zeros(B, 2, 32, 43).pin_memory() for _ in range(nb)]
    xd = [torch.empty(B, 6, 32, 43, device=dev()) for _ in range(2)]
    yd = [torch.empty(B, 2, 32, 43, device=dev()) for _ in range(2)]
    for chunk in (B, 2):
        plan.set_chunk(chunk)
        for y in yh:
            y.zero_()
        for k in range(nb):
            plan.forward_host_async(xh[k], None, pp, yh[k], xd[k % 2], yd[k % 2], packed, slot=k % 2)
        plan.host_wait(0)
        plan.host_wait(1)     # downloads are ordered on one copy stream: the last two waits cover all six
        for k in range(nb):
            assert torch.equal(yh[k], refs[k]), (chunk, k)
    plan.set_chunk(B)
    with pytest.raises(RuntimeError):
        plan.host_wait(3)     # nothing in flight on that slot
    with pytest.raises(RuntimeError):
        plan.forward_host_async(xh[0], None, pp, yh[0], xd[0], yd[0], packed, slot=9)


@pytest.mark.parametrize("cin,cin1,cout,h,w,b,pool", [(64, 0, 64, 19, 23, 2, False), (64, 0, 64, 40, 53, 2, True),
                                                       (64, 0, 128, 33, 20, 1, False), (128, 0, 128, 21, 27, 2, True),
                                                       (256, 0, 256, 16, 24, 3, False), (64, 64, 64, 20, 26, 2, False),
                                                       (128, 128, 128, 10, 13, 1, False), (64, 0, 64, 16, 8, 1, True)])
def test_conv3x3_halo_kernel(cin, cin1, cout, h, w, b, pool):
    """Halo-resident conv3x3 (csrc/conv_halo.cuh): shifted-descriptor tap views, resident / streamed weights,
    two-source virtual concat, shuffle max-pool."""
    from gelslim_depth_b200.engine import conv3x3_halo_op
    g = torch.Generator().manual_seed(cin + cout + h)
    ct = cin + cin1
    x = bf16r(torch.randn(b, ct, h, w, generator=g))
    wt = bf16r(torch.randn(cout, ct, 3, 3, generator=g) * (2.0 / (9 * ct)) ** 0.5)
    sc, sh = 0.5 + torch.rand(cout, generator=g), 0.3 * torch.randn(cout, generator=g)
    ref = torch.relu(F.conv2d(x, wt, padding=1) * sc[None, :, None, None] + sh[None, :, None, None])
    d = dev()
    xs = nhwc(x).to(torch.bfloat16).to(d)
    s0 = xs[..., :cin].contiguous()
    s1 = xs[..., cin:].contiguous() if cin1 else None
    r = conv3x3_halo_op(s0, pack_w3(wt).to(d), sc.to(d), sh.to(d), relu=True, src1=s1, pool=pool)
    torch.cuda.synchronize()
    out = r[0] if pool else r
    check_close(out.permute(0, 3, 1, 2), ref, "halo conv3x3")
    if pool:
        want = F.max_pool2d(out.permute(0, 3, 1, 2).float().cpu(), 2)
        assert torch.equal(r[1].permute(0, 3, 1, 2).float().cpu(), want)


@pytest.mark.parametrize("cin,cin1,cout,h,w,b,pool,block_n", [
    (64, 0, 128, 33, 20, 1, False, 0),      # 3 x 3 x 1 = 9 tiles: odd number of M groups, the peer CTA of the last pair idles
    (128, 0, 128, 21, 27, 2, True, 0),      # fused max-pool in both CTAs
    (256, 0, 256, 16, 24, 3, False, 256),   # N = 256 UMMA (128 weight rows per CTA)
    (128, 128, 128, 10, 13, 1, False, 0),   # two-source virtual concat, a single tile: one pair, idle peer
    (64, 64, 256, 40, 53, 2, False, 0),     # several N tiles per M group
    (64, 0, 64, 40, 53, 2, True, 0)])       # resident weights, 32 rows per CTA (not selected by default; must still be exact)
def test_conv3x3_halo_kernel_cta_pair(cin, cin1, cout, h, w, b, pool, block_n, monkeypatch):
    """CTA-pair variant of the halo conv (thread-block cluster of 2, tcgen05.mma.cta_group::2, csrc/conv_halo.cuh): the
    launch rule only pairs CTAs when a layer has a full wave of work, so GSD_CTA2=2 forces it on small shapes.  Result
    must match the fp32 reference AND be bit-identical to the single-CTA kernel (same accumulation order)."""
    from gelslim_depth_b200.engine import conv3x3_halo_op
    g = torch.Generator().manual_seed(cin + cout + h)
    ct = cin + cin1
    x = bf16r(torch.randn(b, ct, h, w, generator=g))
    wt = bf16r(torch.randn(cout, ct, 3, 3, generator=g) * (2.0 / (9 * ct)) ** 0.5)
    sc, sh = 0.5 + torch.rand(cout, generator=g), 0.3 * torch.randn(cout, generator=g)
    ref = torch.relu(F.conv2d(x, wt, padding=1) * sc[None, :, None, None] + sh[None, :, None, None])
    d = dev()
    xs = nhwc(x).to(torch.bfloat16).to(d)
    s0 = xs[..., :cin].contiguous()
    s1 = xs[..., cin:].contiguous() if cin1 else None
    args = (s0, pack_w3(wt).to(d), sc.to(d), sh.to(d))
    results = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("GSD_CTA2", mode)
        r = conv3x3_halo_op(*args, relu=True, src1=s1, pool=pool, block_n=block_n)
        torch.cuda.synchronize()
        results[mode] = r if pool else (r,)
    check_close(results["2"][0].permute(0, 3, 1, 2), ref, "halo conv3x3, CTA pair")
    for single, pair in zip(results["0"], results["2"]):
        assert torch.equal(single, pair)


@pytest.mark.parametrize("cin,cin1,cout,h,w,b,pool,bn", [
    (128, 0, 256, 20, 26, 3, False, 256),   # 20x26 bottleneck shape, N = 256: 128 weight rows per CTA
    (64, 64, 128, 10, 13, 1, False, 128),   # virtual concat with padding, 2 tiles = one pair
    (64, 0, 128, 21, 27, 1, True, 128),     # fused max-pool; 6 tiles... odd counts leave a duplicate peer tile
    (128, 0, 128, 17, 9, 3, False, 128)])   # 3 x 1 x 3 = 9 tiles: the last pair's peer is a duplicate
def test_conv_tc_kernel_cta_pair(cin, cin1, cout, h, w, b, pool, bn, monkeypatch):
    """CTA-pair variant of the tap-streaming conv (csrc/conv_tc.cuh, cta_group::2): bit-identical to single CTAs."""
    from gelslim_depth_b200.engine import conv_op
    g = torch.Generator().manual_seed(cin + cout + h)
    ct = cin + cin1
    x = bf16r(torch.randn(b, ct, h, w, generator=g))
    wt = bf16r(torch.randn(cout, ct, 3, 3, generator=g) * (2.0 / (9 * ct)) ** 0.5)
    sc, sh = 0.5 + torch.rand(cout, generator=g), 0.3 * torch.randn(cout, generator=g)
    ref = torch.relu(F.conv2d(x, wt, padding=1) * sc[None, :, None, None] + sh[None, :, None, None])
    d = dev()
    xs = nhwc(x).to(torch.bfloat16).to(d)
    s0 = xs[..., :cin].contiguous()
    s1 = xs[..., cin:].contiguous() if cin1 else None
    results = {}
    for mode in ("0", "2"):
        monkeypatch.setenv("GSD_CTA2", mode)
        r = conv_op(s0, pack_w3(wt).to(d), sc.to(d), sh.to(d), TAPS3, relu=True, src1=s1, pool=pool, block_n=bn)
        torch.cuda.synchronize()
        results[mode] = r if pool else (r,)
    check_close(results["2"][0].permute(0, 3, 1, 2), ref, "tap-streaming conv3x3, CTA pair")
    for single, pair in zip(results["0"], results["2"]):
        assert torch.equal(single, pair)


def test_first_layer_halo_kernel():
    """inc.double_conv.0 through the SW32 / 16-channel variant of the halo kernel."""
    from gelslim_depth_b200.engine import conv3x3_halo_op
    g = torch.Generator().manual_seed(11)
    for cin, h, w in ((3, 21, 37), (6, 40, 53), (6, 16, 8)):
        x = bf16r(torch.rand(2, cin, h, w, generator=g))
        wt = bf16r(torch.randn(64, cin, 3, 3, generator=g) * 0.3)
        sc, sh = 0.5 + torch.rand(64, generator=g), 0.3 * torch.randn(64, generator=g)
        ref = torch.relu(F.conv2d(x, wt, padding=1) * sc[None, :, None, None] + sh[None, :, None, None])
        x16 = torch.zeros(2, h, w, 16)
        x16[..., :cin] = nhwc(x)
        d = dev()
        out = conv3x3_halo_op(x16.to(torch.bfloat16).to(d), pack_w3(wt, 16).to(d), sc.to(d), sh.to(d), relu=True)
        torch.cuda.synchronize()
        check_close(out.permute(0, 3, 1, 2), ref, f"first layer (halo) cin={cin}")


def test_fp32_parity_mode_within_1e3_mm(golden_full):
    """north star: max-abs depth error <= 1e-3 mm in fp32.  1 network unit = (max-min)/norm_scale = 2.131 mm with
    the shipped normalisation (config_unet_bigdata.py:42-43), so the bound is 4.7e-4 network units; the fp32
    FFMA path is compared with the reference-generated golden output and per-layer with the fp32/fp64 oracle."""
    from gelslim_depth_b200.processing_utils.complete_prediction import predict_depth_from_RGB
    mm_per_unit = 1.9180814027786255 / 0.9
    for tag in ("g2_eval", "g3_eval"):
        g = golden_full[tag]
        net, sd = build_net(g, dev())
        net.set_precision("fp32")
        y = net(x=g["x"].to(dev())).cpu()
        err_mm = float((y - g["y"]).abs().max()) * mm_per_unit
        assert float(g["y"].abs().max()) > 0.5, "non-degenerate output required"
        assert err_mm <= 1e-3, f"{tag}: max-abs depth error {err_mm:.3e} mm"
        y64 = oracle.unet_forward(sd, g["x"], dtype=torch.float64)
        assert rel_l2(y, y64) < 1e-5
    # whole pipeline in mm (G3: area down/up-sampling + normalisation) vs the oracle
    torch.manual_seed(4)
    from gelslim_depth_b200.models.unet import UNet
    net = UNet(3, 1)
    sd = oracle.conditioned_state_dict(net.state_dict(), seed=8)
    net.load_state_dict(sd)
    net = net.to(dev()).eval().set_precision("fp32")
    gen = torch.Generator().manual_seed(6)
    raw = torch.randint(0, 256, (2, 6, 64, 85), generator=gen).float()
    base = torch.randint(0, 256, (1, 6, 64, 85), generator=gen).float()
    fingers = oracle.split_fingers(oracle.get_difference_image(raw, base))
    cfg = shipped_cfg((32, 43))
    ref = oracle.predict_depth_from_RGB(fingers, lambda t: oracle.unet_forward(sd, t), (64, 85), cfg)
    got = predict_depth_from_RGB(fingers.to(dev()), net, (64, 85), cfg).cpu()
    assert float((got - ref).abs().max()) <= 1e-3, float((got - ref).abs().max())


@pytest.mark.parametrize("cin,cin1,cout,h,w,b", [(64, 0, 64, 19, 23, 2), (64, 0, 128, 16, 8, 1), (128, 0, 256, 21, 27, 2),
                                                 (64, 64, 64, 20, 26, 2), (256, 0, 128, 10, 13, 3),
                                                 (64, 0, 64, 32, 16, 1), (128, 0, 64, 16, 9, 2), (64, 0, 64, 1, 1, 1),
                                                 (128, 0, 256, 20, 26, 2), (64, 64, 128, 40, 53, 1), (64, 0, 64, 8, 16, 2),
                                                 (128, 0, 64, 7, 30, 1)])   # the last four select the 16 x 8 pixel tile
def test_wgrad3x3_tcgen05(cin, cin1, cout, h, w, b):
    """conv weight gradient = GEMM over pixels with MN-major operands (csrc/wgrad_tc.cuh) vs autograd."""
    from gelslim_depth_b200.engine import wgrad3x3_op
    g = torch.Generator().manual_seed(cin + cout + h)
    ct = cin + cin1
    x = bf16r(torch.randn(b, ct, h, w, generator=g))
    dz = bf16r(torch.randn(b, cout, h, w, generator=g))
    wt = torch.zeros(cout, ct, 3, 3, requires_grad=True)
    F.conv2d(x, wt, padding=1).backward(dz)
    ref = wt.grad.reshape(cout, ct, 9).permute(0, 2, 1)        # [co][tap][ci]
    d = dev()
    xs = nhwc(x).to(torch.bfloat16).to(d)
    s0 = xs[..., :cin].contiguous()
    s1 = xs[..., cin:].contiguous() if cin1 else None
    got = wgrad3x3_op(s0, nhwc(dz).to(torch.bfloat16).to(d), x1=s1).cpu()
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) < 2e-3 * scale + 1e-3, (float((got - ref).abs().max()), scale)
    assert rel_l2(got, ref) < 1e-3


def test_frame_pairs_split_and_uint8_ingest():
    """Left/Right split (general_dataset.py:71), difference image, area down-sampling, normalisation
    fused into the first kernel, from float and from uint8 camera frames."""
    from gelslim_depth_b200.models.unet import UNet
    from gelslim_depth_b200.processing_utils.complete_prediction import predict_depth_from_frame_pairs
    torch.manual_seed(4)
    net = UNet(3, 1)
    sd = oracle.conditioned_state_dict(net.state_dict(), seed=8)
    net.load_state_dict(sd)
    net = net.to(dev()).eval()
    g = torch.Generator().manual_seed(12)
    raw8 = torch.randint(0, 256, (3, 6, 64, 85), generator=g, dtype=torch.uint8)
    base = torch.randint(0, 256, (1, 6, 64, 85), generator=g).float()
    cfg = shipped_cfg((32, 43))
    fingers = oracle.split_fingers(oracle.get_difference_image(raw8.float(), base))           # (6, 3, 64, 85)
    ref = oracle.predict_depth_from_RGB(fingers, lambda t: oracle.unet_forward(sd, t), (64, 85), cfg)
    ref = ref.view(2, 3, 64, 85).permute(1, 0, 2, 3)
    got_f = predict_depth_from_frame_pairs(raw8.float().to(dev()), base.to(dev()), net, (64, 85), cfg)
    got_u = predict_depth_from_frame_pairs(raw8.to(dev()), base.to(dev()), net, (64, 85), cfg)
    assert got_f.shape == (3, 2, 64, 85)
    assert torch.equal(got_f, got_u)                       # uint8 ingest is bit-identical to float frames
    assert rel_l2(got_f, ref) < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["hwc_u8", "chw_u8", "chw_f32"])
def test_depth_stream_ring_buffer_graph_replay(layout):
    """camera frames -> pinned ring slot -> one CUDA-graph replay -> depth map; identical to the batched
    entry point on the same frames, for interleaved uint8 camera frames and the two planar layouts."""
    from gelslim_depth_b200.models.unet import UNet
    from gelslim_depth_b200.processing_utils.complete_prediction import predict_depth_from_frame_pairs
    from gelslim_depth_b200.streaming import DepthStream
    torch.manual_seed(5)
    net = UNet(3, 1)
    sd = oracle.conditioned_state_dict(net.state_dict(), seed=9)
    net.load_state_dict(sd)
    net = net.to(dev()).eval()
    g = torch.Generator().manual_seed(13)
    raw8 = torch.randint(0, 256, (6, 6, 64, 85), generator=g, dtype=torch.uint8)
    base = torch.randint(0, 256, (6, 64, 85), generator=g).float()
    cfg = shipped_cfg((32, 43))
    want = predict_depth_from_frame_pairs(raw8.to(dev()), base[None].to(dev()), net, (64, 85), cfg).cpu()
    fingers = oracle.split_fingers(oracle.get_difference_image(raw8[:1].float(), base[None]))
    ref0 = oracle.predict_depth_from_RGB(fingers, lambda t: oracle.unet_forward(sd, t), (64, 85), cfg)[:, 0]
    stream = DepthStream(net, cfg, (64, 85), base_tactile_image=base, output_size=(64, 85), layout=layout, frame_pairs=True, slots=3)
    frames = {"hwc_u8": raw8.permute(0, 2, 3, 1).contiguous(), "chw_u8": raw8, "chw_f32": raw8.float()}[layout]
    # two frames in flight, then one at a time round the ring (6 frames over 3 slots: every slot is reused)
    t0, t1 = stream.push(frames[0]), stream.push(frames[1])
    got = [stream.result(t0).clone(), stream.result(t1).clone()]
    for i in range(2, 6):
        got.append(stream(frames[i]).clone())
    got = torch.stack(got)
    assert got.shape == (6, 2, 64, 85)
    assert torch.equal(got, want)
    assert rel_l2(got[0], ref0) < 2e-2
    assert len(stream.latencies_ms) == 6
    with pytest.raises(ValueError):
        stream.push(torch.zeros(3, 3, 3))
    # zero-copy ingest: the producer writes into the pinned slot itself
    ticket, buf = stream.acquire()
    buf.copy_(frames[4])
    assert torch.equal(stream.result(stream.submit(ticket)), want[4])


@pytest.mark.gpu
def test_dataset_preprocessing_vs_reference_generaldataset():
    """Left/Right split + difference image + area down-sampling + normalisation of a whole object tensor in
    one launch per tensor (+ the Gaussian depth blur), against samples served by the unmodified reference GeneralDataset."""
    import os
    from gelslim_depth_b200.processing_utils.dataset_utils import preprocess_object_tensors
    from gelslim_depth_b200.processing_utils.image_utils import blur_depth_images
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "dataset_preprocess.pt"), weights_only=False)
    data = {k: v.to(dev()) for k, v in fx["data"].items()}
    for name, case in fx["cases"].items():
        kw = case["kwargs"]
        got = preprocess_object_tensors(
            data["tactile_image"], data["depth_image"], data["base_tactile_image"], case["input_tactile_image_size"],
            kw["image_normalization_method"], case["image_normalization_parameters"], kw["depth_normalization_method"],
            case["depth_normalization_parameters"], kw["norm_scale"], separate_fingers=kw["separate_fingers"],
            use_difference_image=kw["use_difference_image"], interp_method=kw["interp_method"],
            depth_image_blur_kernel=kw["depth_image_blur_kernel"])
        assert got["tactile_image"].shape == case["tactile_image"].shape, name
        assert torch.allclose(got["tactile_image"].cpu(), case["tactile_image"], rtol=1e-5, atol=1e-5), name
        assert torch.allclose(got["depth_image"].cpu(), case["depth_image"], rtol=1e-5, atol=2e-6), name
    d = torch.rand(2, 1, 9, 7, generator=torch.Generator().manual_seed(1))
    assert torch.allclose(blur_depth_images(d.to(dev()), 7).cpu(), oracle.blur_depth_images(d, 7), rtol=1e-5, atol=1e-6)
    with pytest.raises(RuntimeError):
        blur_depth_images(d.to(dev()), 4)            # even kernel sizes have no centre tap (the library rejects them)


@pytest.mark.gpu
def test_integration_md_ctypes_stub_runs_as_printed():
    """INTEGRATION.md section B shows the binding a maintainer would add to the reference; the code block is executed
    verbatim here (only the library path is made absolute) and checked against the oracle."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"```python\nimport ctypes as C, torch\n(.*?)```", src, re.S).group(0)
    code = block[len("```python\n"):-3].replace('C.CDLL("libgsd_b200.so")',
                                                'C.CDLL(%r)' % os.path.join(root, "gelslim_depth_b200", "libgsd_b200.so"))
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    from gelslim_depth_b200.models.unet import UNet
    torch.manual_seed(0)
    net = UNet(6, 2)
    sd = oracle.conditioned_state_dict(net.state_dict(), seed=5)
    net.load_state_dict(sd)
    net = net.to(dev()).eval()
    x = torch.rand(2, 6, 48, 59, generator=torch.Generator().manual_seed(1))
    y = ns["unet_forward_b200"](net, x.to(dev()))
    torch.cuda.synchronize()
    assert rel_l2(y, oracle.unet_forward(sd, x)) < 2e-2


@pytest.mark.parametrize("case", ["f32_diff", "u8_nchw", "u8_nhwc", "pairs_u8", "plain"])
def test_first_conv_fused_prologue_bit_identical(case, monkeypatch):
    """north_star bullet 4: get_difference_image (image_utils.py:6-10), the Left/Right split (general_dataset.py:71)
    and normalize_tactile_image (normalization_utils.py:29-34) run inside the first conv's producer warps
    (csrc/conv_first.cuh): no prologue launch, no normalised tensor in HBM.  Must equal -- bit for bit -- the two-pass
    form (prologue_kernel + TMA-fed first conv, forced by GSD_NO_FUSED_PROLOGUE) on ragged geometries, and the oracle
    within the bf16 bound."""
    from gelslim_depth_b200.engine import make_prepost
    from gelslim_depth_b200.models.unet import UNet
    pairs = case == "pairs_u8"
    cin = 3 if pairs else 6
    torch.manual_seed(21)
    net = UNet(cin, 1 if pairs else 2, layer_dimensions=[64, 128, 256])
    sd = oracle.conditioned_state_dict(net.state_dict(), seed=5)
    net.load_state_dict(sd)
    net = net.to(dev()).eval()
    g = torch.Generator().manual_seed(33)
    B, H, W = 3, 40, 53                                  # ragged 16 x 8 tiles, still halo-kernel territory (>= 75 % useful rows)
    raw8 = torch.randint(0, 256, (B, 6, H, W), generator=g, dtype=torch.uint8)
    base = torch.randint(0, 256, (1 if case != "f32_diff" else B, 6, H, W), generator=g).float()
    scale, shift = [1 / 255.0, 1 / 200.0, 1 / 255.0], [0.0, 0.25, -0.5]
    kw = dict(use_diff=case != "plain", base_batch=base.shape[0], in_scale=scale, in_shift=shift, out_scale=-2.1, out_shift=0.3,
              split_fingers=pairs)
    if case == "u8_nhwc":
        x = raw8.permute(0, 2, 3, 1).contiguous().to(dev())
        pp = make_prepost(cin, (H, W), (H, W), input_u8=2, **kw)
    elif case in ("u8_nchw", "pairs_u8"):
        x = raw8.to(dev())
        pp = make_prepost(cin, (H, W), (H, W), input_u8=1, **kw)
    else:
        x = raw8.float().to(dev())
        pp = make_prepost(cin, (H, W), (H, W), **kw)
    n_net = 2 * B if pairs else B
    plan = net.plan_for(n_net, H, W, dev())
    packed = net.packed_weights(plan)
    based = base.to(dev()) if case != "plain" else None

    def run():
        y = torch.empty(n_net, net.n_classes, H, W, device=dev())
        plan.forward(x, based, pp, y, packed)
        torch.cuda.synchronize()
        return y

    y_fused = run()
    assert plan.first_fused
    n_fused = plan.launches
    monkeypatch.setenv("GSD_NO_FUSED_PROLOGUE", "1")
    y_two_pass = run()
    assert not plan.first_fused and plan.launches == n_fused + 1         # the prologue pass is the only extra launch
    monkeypatch.delenv("GSD_NO_FUSED_PROLOGUE")
    assert torch.equal(y_fused, y_two_pass)
    # oracle: the reference's own order of operations on the CPU
    xin = raw8.float()
    if case != "plain":
        xin = oracle.get_difference_image(xin, base)
    if pairs:
        xin = oracle.split_fingers(xin)
    sc = torch.tensor([scale[min(c, 2)] for c in range(cin)]).view(1, cin, 1, 1)
    sh = torch.tensor([shift[min(c, 2)] for c in range(cin)]).view(1, cin, 1, 1)
    ref = oracle.unet_forward(sd, xin * sc + sh) * -2.1 + 0.3
    assert rel_l2(y_fused, ref) < 2e-2
    # Both forms above add the 64-channel layers' BatchNorm shift with one extra UMMA per tile (csrc/bias_mma.cuh: bf16
    # hi + lo split of the fp32 constant, exact products, fp32 accumulation).  GSD_NO_BIAS_MMA (read when a plan binds)
    # restores the epilogue add: same mathematics, different rounding order -- a few outputs move by one bf16 ulp per layer.
    monkeypatch.setenv("GSD_NO_BIAS_MMA", "1")
    net._plans.clear()
    plan = net.plan_for(n_net, H, W, dev())
    y_epi = run()
    assert not plan.first_fused                      # the fused kernel has no epilogue-constant path
    assert rel_l2(y_epi, y_fused) < 2e-2, rel_l2(y_epi, y_fused)      # measured 3e-3 .. 8e-3 (11 layers of one-ulp flips)
    assert rel_l2(y_epi, ref) < 2e-2
