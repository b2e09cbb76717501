"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares; the drop-in
module keeps the reference's interface; nothing falls back to CPU arithmetic."""
import ctypes
import os
import re
import types

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gsd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gelslim_depth_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 15
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/gsd_b200.h but not exported"
        assert s in _lib.SYMBOLS, f"{s} has no ctypes prototype in _lib.py"
    assert _lib.lib.gsd_abi_version() == 1


def test_no_gpu_means_loud_failure():
    from gelslim_depth_b200 import _lib
    from gelslim_depth_b200.models.unet import UNet
    net = UNet(3, 1).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x=torch.zeros(1, 3, 32, 43))
    if _lib.lib.gsd_device_count() == 0:
        g = _lib.Geometry()
        g.batch, g.in_channels, g.height, g.width, g.n_classes, g.n_dims = 1, 3, 32, 43, 1, 5
        for i, d in enumerate([64, 128, 256, 512, 1024]):
            g.dims[i] = d
        h = ctypes.c_void_p()
        rc = _lib.lib.gsd_plan_create(ctypes.byref(h), ctypes.byref(g), 0)
        assert rc != 0 and _lib.lib.gsd_last_error()


def test_state_dict_layout_matches_reference(golden_small, golden_full):
    from gelslim_depth_b200.models.unet import UNet
    net = UNet(3, 1)
    keys = list(net.state_dict().keys())
    assert len(keys) == 118                                   # 64 params + 54 BN buffers
    assert len(list(net.parameters())) == 64
    assert keys[0] == "inc.double_conv.0.weight" and keys[-1] == "outc.conv.bias"
    assert "down.3.maxpool_conv.1.double_conv.4.num_batches_tracked" in keys
    assert net.state_dict()["up.0.up.weight"].shape == (1024, 512, 2, 2)
    assert net.state_dict()["outc.conv.weight"].shape == (1, 64, 1, 1)
    # same default initialisation as the reference under the same seed (digest recorded by make_golden.py)
    import oracle
    g = golden_full["g2_eval"]
    torch.manual_seed(g["module_seed"])
    net6 = UNet(6, 2)
    assert oracle.state_dict_digest(oracle.conditioned_state_dict(net6.state_dict(), g["init_seed"])) == g["digest"]
    assert (net6.n_channels, net6.n_classes, net6.bilinear) == (6, 2, False)


def test_unsupported_configs_are_rejected():
    from gelslim_depth_b200.models.unet import UNet
    for kw in (dict(layer_dimensions=[4, 8, 16]), dict(kernel_size=5), dict(layer_dimensions=[64, 256, 1024]),
               dict(maxpool_size=3)):
        with pytest.raises(NotImplementedError):
            UNet(3, 1, **kw)


def test_normalisation_constants_match_reference_tables(golden_processing):
    from gelslim_depth_b200.processing_utils.normalization_utils import image_affine_constants, depth_affine_constants
    p4 = ([1.0, 2.0, 3.0], [200.0, 210.0, 220.0], [100.0, 110.0, 120.0], [50.0, 60.0, 70.0])
    x = golden_processing["norm_in"]
    for m in ("mean_std", "0_255_to_-1_1", "0_255_to_0_1"):
        s, t = image_affine_constants(m, 0.9, p4)
        got = torch.stack([x[:, c] * s[min(c, len(s) - 1)] + t[min(c, len(t) - 1)] for c in range(3)], dim=1)
        assert torch.allclose(got, golden_processing["norm_img_" + m], rtol=1e-5, atol=1e-5), m
    with pytest.raises(TypeError):
        image_affine_constants("min_max_to_-1_1", 0.9, p4)       # reference raises too (recorded in the fixture)
    assert golden_processing["norm_img_min_max_to_-1_1_raises"] == "TypeError"
    d = golden_processing["depth_in"]
    dp = (-1.9180814027786255, 0.0, -0.4, 0.3)
    for m in ("min_max_to_-1_1", "mean_std", "min_max_to_0_1", "min_max_to_0_-1"):
        scale, bias, den = depth_affine_constants(m, 0.9, dp)
        assert torch.allclose(d * (den / scale) + bias, golden_processing["denorm_depth_" + m], rtol=1e-5, atol=1e-6)
        assert torch.allclose(d * (scale / den) - scale * bias / den, golden_processing["norm_depth_" + m], rtol=1e-5, atol=1e-6)


def test_training_plan_layouts_bf16_and_fp32_without_a_gpu():
    """gsd_debug_train_plan_create (no GPU): the bf16 plan and the fp32 parity plan (geometry.dtype = GSD_DTYPE_FP32,
    csrc/train_plan_f32.h) describe the same parameters in the same order; the fp32 workspace (fp32 activations and operand
    copies, no split-K weight-gradient accumulators) is larger than the bf16 one but less than 2.5x; other dtypes are rejected."""
    import ctypes as C
    from gelslim_depth_b200 import _lib
    from gelslim_depth_b200._lib import lib

    def plan(dtype):
        g = _lib.Geometry()
        g.batch, g.in_channels, g.height, g.width, g.n_classes, g.n_dims = 2, 6, 40, 53, 2, 5
        for i, d in enumerate((64, 128, 256, 512, 1024)):
            g.dims[i] = d
        g.dtype, g.mode = dtype, _lib.MODE_TRAIN
        h = C.c_void_p()
        rc = lib.gsd_debug_train_plan_create(C.byref(h), C.byref(g))
        return rc, h

    rc, hb = plan(_lib.DTYPE_BF16)
    assert rc == 0, lib.gsd_last_error()
    rc, hf = plan(_lib.DTYPE_FP32)
    assert rc == 0, lib.gsd_last_error()
    n = lib.gsd_train_plan_num_params(hb)
    assert n == lib.gsd_train_plan_num_params(hf) == 18 * 3 + 4 * 2 + 2
    assert lib.gsd_train_plan_num_bn(hb) == lib.gsd_train_plan_num_bn(hf) == 18
    a, b = (C.c_longlong * n)(), (C.c_longlong * n)()
    assert lib.gsd_train_plan_param_numel(hb, a, n) == n and lib.gsd_train_plan_param_numel(hf, b, n) == n
    assert list(a) == list(b) and a[0] == 64 * 6 * 9
    wb, wf = lib.gsd_train_plan_workspace_bytes(hb), lib.gsd_train_plan_workspace_bytes(hf)
    assert wb < wf < 2.5 * wb, (wb, wf)
    lib.gsd_train_plan_destroy(hb)
    lib.gsd_train_plan_destroy(hf)
    rc, _ = plan(7)
    assert rc != 0 and b"dtype" in lib.gsd_last_error()
