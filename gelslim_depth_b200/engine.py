"""Host-side driver of libgsd_b200.so: plan cache, torch-owned workspaces, packed-weight cache.

PyTorch is the allocator and stream provider only; every FLOP happens inside the library.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import lib, check


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def make_prepost(in_channels: int, raw_hw: Tuple[int, int], out_hw: Tuple[int, int], use_diff: bool = False,
                 base_batch: int = 1, in_scale: Sequence[float] = (1.0,), in_shift: Sequence[float] = (0.0,),
                 out_scale: float = 1.0, out_shift: float = 0.0, split_fingers: bool = False,
                 input_u8: bool = False) -> _lib.PrePost:
    pp = _lib.PrePost()
    pp.use_diff = int(use_diff)
    pp.base_batch = int(base_batch)
    pp.raw_height, pp.raw_width = int(raw_hw[0]), int(raw_hw[1])
    pp.out_height, pp.out_width = int(out_hw[0]), int(out_hw[1])
    for c in range(8):
        pp.in_scale[c] = float(in_scale[min(c, len(in_scale) - 1)])
        pp.in_shift[c] = float(in_shift[min(c, len(in_shift) - 1)])
    pp.out_scale, pp.out_shift = float(out_scale), float(out_shift)
    pp.split_fingers, pp.input_u8 = int(split_fingers), int(input_u8)
    return pp


class Plan:
    """One gsd_plan + the torch tensors that back its workspace."""

    def __init__(self, batch, in_channels, height, width, n_classes, dims, device: torch.device,
                 dtype=_lib.DTYPE_BF16, mode=_lib.MODE_INFER):
        g = _lib.Geometry()
        g.batch, g.in_channels, g.height, g.width, g.n_classes = batch, in_channels, height, width, n_classes
        g.n_dims = len(dims)
        for i, d in enumerate(dims):
            g.dims[i] = int(d)
        g.dtype, g.mode = dtype, mode
        self.geometry = g
        self.device = device
        self.handle = C.c_void_p()
        check(lib.gsd_plan_create(C.byref(self.handle), C.byref(g), device.index or 0), "gsd_plan_create")
        self.workspace = torch.empty(lib.gsd_plan_workspace_bytes(self.handle), dtype=torch.uint8, device=device)
        self.packed_bytes = lib.gsd_plan_packed_bytes(self.handle)

    def set_chunk(self, frames: int, first: int = 0, last: int = 0):
        check(lib.gsd_plan_set_chunk(self.handle, int(frames)), "gsd_plan_set_chunk")
        if first or last:
            check(lib.gsd_plan_set_chunk_ramp(self.handle, int(first), int(last)), "gsd_plan_set_chunk_ramp")

    @property
    def launches(self) -> int:
        return lib.gsd_plan_forward_launches(self.handle)

    @property
    def first_fused(self) -> bool:
        """the last forward ran the input prologue inside the first conv (no prologue launch, no 16-channel tensor)"""
        return bool(lib.gsd_plan_first_fused(self.handle))

    @property
    def conv_flops(self) -> float:
        return lib.gsd_plan_conv_flops(self.handle)

    def pack(self, params, bn_buffers, packed: torch.Tensor, state: Optional[torch.Tensor] = None):
        """gsd_pack_weights; with `state` (4 x int64 device tensor, zeroed when `packed` is new) the re-pack happens only if
        the device-side content fingerprint of the parameters changed (gsd_pack_weights_if_changed)."""
        n_p, n_b = lib.gsd_plan_num_params(self.handle), lib.gsd_plan_num_bn_buffers(self.handle)
        if len(params) != n_p or len(bn_buffers) != n_b:
            raise ValueError(f"expected {n_p} params / {n_b} BN buffers, got {len(params)} / {len(bn_buffers)}")
        for t in list(params) + list(bn_buffers):
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("parameters must be contiguous fp32 tensors on the plan's device")
        pa = (C.c_void_p * n_p)(*[t.data_ptr() for t in params])
        ba = (C.c_void_p * n_b)(*[t.data_ptr() for t in bn_buffers])
        if state is None:
            check(lib.gsd_pack_weights(self.handle, pa, ba, _ptr(packed), _stream(self.device)), "gsd_pack_weights")
        else:
            check(lib.gsd_pack_weights_if_changed(self.handle, pa, ba, _ptr(packed), _ptr(state), _stream(self.device)),
                  "gsd_pack_weights_if_changed")

    def forward(self, x, base, pp, y, packed):
        check(lib.gsd_forward(self.handle, _ptr(x), _ptr(base), C.byref(pp), _ptr(y), _ptr(self.workspace),
                              _ptr(packed), _stream(self.device)), "gsd_forward")

    def forward_profiled(self, x, base, pp, y, packed):
        """-> list of (ms, flops) per launch in network order (prologue, convs..., head)."""
        cap = 64
        ms = (C.c_float * cap)()
        fl = (C.c_double * cap)()
        n = C.c_int(0)
        check(lib.gsd_forward_profiled(self.handle, _ptr(x), _ptr(base), C.byref(pp), _ptr(y), _ptr(self.workspace),
                                       _ptr(packed), _stream(self.device), ms, fl, cap, C.byref(n)),
              "gsd_forward_profiled")
        return [(ms[i], fl[i]) for i in range(n.value)]

    def activations(self):
        """per-layer taps of the last forward (parity tests): {reference hook name: fp32 NCHW tensor}.  Names follow
        forward hooks on the reference module (unet.py:7-57): '<prefix>.double_conv.2' / '.5' = post-ReLU outputs,
        'up.i.up' = transposed-conv output."""
        depth = self.geometry.n_dims - 1
        names = []
        for l in range(depth + 1):
            prefix = "inc" if l == 0 else f"down.{l - 1}.maxpool_conv.1"
            names += [f"{prefix}.double_conv.2", f"{prefix}.double_conv.5"]
        for i in range(depth):
            names += [f"up.{i}.up", f"up.{i}.conv.double_conv.2", f"up.{i}.conv.double_conv.5"]
        n = lib.gsd_debug_num_activations(self.handle)
        assert n == len(names)
        out = {}
        for idx, name in enumerate(names):
            c, h, w = C.c_int(), C.c_int(), C.c_int()
            check(lib.gsd_debug_activation_shape(self.handle, idx, C.byref(c), C.byref(h), C.byref(w)), "gsd_debug_activation_shape")
            t = torch.empty(self.geometry.batch, c.value, h.value, w.value, dtype=torch.float32, device=self.device)
            rc = lib.gsd_debug_read_activation(self.handle, idx, _ptr(self.workspace), _ptr(t), _stream(self.device))
            if rc != 0 and idx == n - 1:
                continue                     # fused 1x1 head: the last unit's output is never stored
            check(rc, "gsd_debug_read_activation")
            out[name] = t
        return out

    def forward_host(self, x_host, base, pp, y_host, x_dev, y_dev, packed):
        check(lib.gsd_forward_host(self.handle, _ptr(x_host), _ptr(base), C.byref(pp), _ptr(y_host), _ptr(x_dev),
                                   _ptr(y_dev), _ptr(self.workspace), _ptr(packed), _stream(self.device)),
              "gsd_forward_host")

    def forward_host_async(self, x_host, base, pp, y_host, x_dev, y_dev, packed, slot: int):
        """Enqueue upload | compute | download on staging set `slot` and return (gsd_forward_host_async); consecutive
        calls on rotating slots overlap.  `host_wait(slot)` blocks until that slot's y_host is complete."""
        check(lib.gsd_forward_host_async(self.handle, _ptr(x_host), _ptr(base), C.byref(pp), _ptr(y_host), _ptr(x_dev),
                                         _ptr(y_dev), _ptr(self.workspace), _ptr(packed), _stream(self.device),
                                         int(slot)), "gsd_forward_host_async")

    def host_wait(self, slot: int):
        check(lib.gsd_forward_host_wait(self.handle, int(slot)), "gsd_forward_host_wait")

    def __del__(self):
        try:
            if self.handle:
                lib.gsd_plan_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


class PlanCache:
    def __init__(self, capacity: int = 4):
        self.capacity = capacity
        self._plans: "OrderedDict[tuple, Plan]" = OrderedDict()

    def get(self, key, factory) -> Plan:
        if key in self._plans:
            self._plans.move_to_end(key)
            return self._plans[key]
        plan = factory()
        self._plans[key] = plan
        while len(self._plans) > self.capacity:
            self._plans.popitem(last=False)
        return plan

    def clear(self):
        self._plans.clear()


def conv_op(src0, w, scale, shift, taps, relu=True, src1=None, off=(0, 0), cout=None, groups=1, pool=False,
            block_n=0):
    """Single tcgen05 conv (gsd_op_conv_bf16) on NHWC bf16 torch tensors -- used by the parity tests."""
    B, H, W, C0 = src0.shape
    dev = src0.device
    ntaps = len(taps)
    dy = (C.c_int8 * ntaps)(*[t[0] for t in taps])
    dx = (C.c_int8 * ntaps)(*[t[1] for t in taps])
    cout = cout if cout is not None else w.shape[0] // groups
    if groups == 1:
        out = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device=dev)
    else:
        out = torch.empty(B, 2 * H, 2 * W, cout, dtype=torch.bfloat16, device=dev)
    pooled = torch.empty(B, H // 2, W // 2, cout, dtype=torch.bfloat16, device=dev) if pool else None
    C1 = H1 = W1 = 0
    if src1 is not None:
        _, H1, W1, C1 = src1.shape
    check(lib.gsd_op_conv_bf16(_ptr(src0), C0, _ptr(src1), C1, H1, W1, off[0], off[1], B, H, W, _ptr(w), cout, ntaps,
                               C.cast(dy, C.c_void_p), C.cast(dx, C.c_void_p), groups, _ptr(scale), _ptr(shift),
                               int(relu), _ptr(out), _ptr(pooled), block_n, dev.index or 0, _stream(dev)),
          "gsd_op_conv_bf16")
    return (out, pooled) if pool else out


def conv3x3_halo_op(src0, w, scale, shift, relu=True, src1=None, off=(0, 0), pool=False, block_n=0, base_off_mode=0):
    """conv3x3 through the halo-resident kernel (gsd_op_conv3x3_halo_bf16); NHWC bf16 torch tensors."""
    B, H, W, C0 = src0.shape
    dev = src0.device
    cout = w.shape[0]
    out = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device=dev)
    pooled = torch.empty(B, H // 2, W // 2, cout, dtype=torch.bfloat16, device=dev) if pool else None
    C1 = H1 = W1 = 0
    if src1 is not None:
        _, H1, W1, C1 = src1.shape
    check(lib.gsd_op_conv3x3_halo_bf16(_ptr(src0), C0, _ptr(src1), C1, H1, W1, off[0], off[1], B, H, W, _ptr(w), cout,
                                       _ptr(scale), _ptr(shift), int(relu), _ptr(out), _ptr(pooled), block_n,
                                       base_off_mode, dev.index or 0, _stream(dev)), "gsd_op_conv3x3_halo_bf16")
    return (out, pooled) if pool else out


def wgrad3x3_op(x0, dz, x1=None, off=(0, 0)):
    """conv3x3 weight gradient (gsd_op_wgrad3x3_bf16): NHWC bf16 x / dz -> fp32 [Cout][9][C0+C1]."""
    B, H, W, C0 = x0.shape
    cout = dz.shape[-1]
    dev = x0.device
    C1 = H1 = W1 = 0
    if x1 is not None:
        _, H1, W1, C1 = x1.shape
    dw = torch.zeros(cout, 9, C0 + C1, dtype=torch.float32, device=dev)
    check(lib.gsd_op_wgrad3x3_bf16(_ptr(x0), C0, _ptr(x1), C1, H1, W1, off[0], off[1], _ptr(dz), cout, B, H, W, _ptr(dw),
                                   dev.index or 0, _stream(dev)), "gsd_op_wgrad3x3_bf16")
    return dw
