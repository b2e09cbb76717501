"""Error behaviour at the C-ABI boundary and host-side cache logic (GPU box)."""
import ctypes as C

import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def test_plan_create_rejects_bad_geometry_with_message():
    from gelslim_depth_b200 import _lib
    g = _lib.Geometry()
    g.batch, g.in_channels, g.height, g.width, g.n_classes, g.n_dims = 1, 3, 32, 43, 1, 3
    for i, d in enumerate([64, 100, 256]):
        g.dims[i] = d
    h = C.c_void_p()
    assert _lib.lib.gsd_plan_create(C.byref(h), C.byref(g), 0) != 0
    assert b"multiple of 64" in _lib.lib.gsd_last_error()
    for i, d in enumerate([64, 128, 256]):
        g.dims[i] = d
    g.height, g.width = 2, 2                                   # too small for 2 poolings
    assert _lib.lib.gsd_plan_create(C.byref(h), C.byref(g), 0) != 0
    assert b"too small" in _lib.lib.gsd_last_error()
    g.height, g.width = 32, 43
    assert _lib.lib.gsd_plan_create(C.byref(h), C.byref(g), 99) != 0       # no such device
    g.dtype = 7
    assert _lib.lib.gsd_plan_create(C.byref(h), C.byref(g), 0) != 0


def test_forward_argument_checks():
    from gelslim_depth_b200._lib import GsdError
    from gelslim_depth_b200.engine import conv_op, make_prepost
    from gelslim_depth_b200.models.unet import UNet
    net = UNet(3, 1, layer_dimensions=[64, 128]).to(dev()).eval()
    x = torch.rand(2, 3, 16, 24, device=dev())
    plan = net.plan_for(2, 16, 24, dev())
    packed = net.packed_weights(plan)
    y = torch.empty(2, 1, 16, 24, device=dev())
    with pytest.raises(GsdError, match="base is NULL"):
        plan.forward(x, None, make_prepost(3, (16, 24), (16, 24), use_diff=True), y, packed)
    with pytest.raises(GsdError, match="smaller than the network input"):
        plan.forward(x, None, make_prepost(3, (8, 8), (16, 24)), y, packed)
    with pytest.raises(ValueError):
        net(x=torch.rand(2, 4, 16, 24, device=dev()))          # wrong channel count
    with pytest.raises(GsdError, match="unsupported input channels"):
        conv_op(torch.zeros(1, 8, 8, 48, dtype=torch.bfloat16, device=dev()), torch.zeros(64, 9 * 48, dtype=torch.bfloat16, device=dev()),
                torch.ones(64, device=dev()), torch.zeros(64, device=dev()), [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1)])


def test_eval_forward_refuses_input_gradients_loudly():
    """the reference's eval forward participates in autograd; ours does not and must say so."""
    from gelslim_depth_b200.models.unet import UNet
    net = UNet(3, 1, layer_dimensions=[64, 128]).to(dev()).eval()
    x = torch.rand(1, 3, 16, 24, device=dev(), requires_grad=True)
    with pytest.raises(NotImplementedError, match="does not propagate gradients"):
        net(x=x)
    with torch.no_grad():
        assert net(x=x).shape == (1, 1, 16, 24)
    assert net(x=x.detach()).requires_grad is False


def test_plan_cache_eviction_and_batch_changes():
    from gelslim_depth_b200.models.unet import UNet
    torch.manual_seed(0)
    net = UNet(3, 1, layer_dimensions=[64, 128])
    sd = oracle.conditioned_state_dict(net.state_dict(), seed=2)
    net.load_state_dict(sd)
    net = net.to(dev()).eval()
    x = torch.rand(7, 3, 24, 31, generator=torch.Generator().manual_seed(1))
    ref = oracle.unet_forward(sd, x)
    full = net(x=x.to(dev())).cpu()
    for b in (1, 2, 3, 5, 6, 7, 4, 1):                          # more distinct shapes than the cache holds
        y = net(x=x[:b].to(dev())).cpu()
        assert torch.allclose(y, full[:b], atol=1e-5), b        # eval-mode results do not depend on the batch
    assert float((full - ref).abs().max()) < 5e-2 * max(1.0, float(ref.abs().max()))


def test_chunked_forward_equals_single_chunk():
    from gelslim_depth_b200.engine import make_prepost
    from gelslim_depth_b200.models.unet import UNet
    torch.manual_seed(0)
    net = UNet(6, 2, layer_dimensions=[64, 128, 256]).to(dev()).eval()
    x = torch.rand(6, 6, 40, 53, device=dev()) * 255
    base = torch.rand(1, 6, 40, 53, device=dev()) * 255
    pp = make_prepost(6, (40, 53), (20, 27), use_diff=True, in_scale=[1 / 255.0], out_scale=-2.0, out_shift=-1.9)
    plan = net.plan_for(6, 20, 27, dev())
    packed = net.packed_weights(plan)
    y1 = torch.empty(6, 2, 20, 27, device=dev())
    plan.forward(x, base, pp, y1, packed)
    plan.set_chunk(4)                                           # 4 + 2 frames
    y2 = torch.empty_like(y1)
    plan.forward(x, base, pp, y2, packed)
    assert torch.equal(y1, y2)
