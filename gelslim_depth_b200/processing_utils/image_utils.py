"""Drop-in for gelslim_depth/processing_utils/image_utils.py (get_difference_image :6-10,
sample_multi_channel_image_to_desired_size :12-15).  Each helper is one launch of the library's
image_affine kernel; inside predict_depth_from_RGB they are fused into the network's prologue
instead.  `blur_depth_images` (:17-19, torchvision gaussian_blur) is one launch of the library's blur kernel."""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import torch

from .._lib import lib, check


def _affine(x: torch.Tensor, size: Tuple[int, int], scale: Sequence[float], shift: Sequence[float],
            base: torch.Tensor = None, split_fingers: bool = False) -> torch.Tensor:
    """out = scale * area_resample(base is None ? x : (x - base + 255)/2) + shift in ONE pass.  split_fingers: x holds
    Left|Right pairs (N, 2C, H, W) and the result is (2N, C, h, w) = cat((x[:, :C], x[:, C:]), 0) (general_dataset.py:71)."""
    if not x.is_cuda:
        raise RuntimeError("gelslim_depth_b200 processing helpers run on a B200 only (no CPU fallback)")
    squeeze = x.dim() == 3
    x4 = (x[None] if squeeze else x).contiguous().float()
    b, c, hr, wr = x4.shape
    if split_fingers:
        if c % 2:
            raise ValueError(f"split_fingers needs an even channel count, got {c}")
        b, c = 2 * b, c // 2
    s8 = (C.c_float * 8)(*[float(scale[min(i, len(scale) - 1)]) for i in range(8)])
    t8 = (C.c_float * 8)(*[float(shift[min(i, len(shift) - 1)]) for i in range(8)])
    base_batch, bptr = 1, C.c_void_p(0)
    if base is not None:
        b4 = (base[None] if base.dim() == 3 else base).contiguous().float().to(x4.device)
        if b4.shape[1:] != x4.shape[1:] or b4.shape[0] not in (1, x4.shape[0]):
            raise ValueError(f"base image shape {tuple(base.shape)} does not broadcast against {tuple(x.shape)}")
        base_batch, bptr = b4.shape[0], C.c_void_p(b4.data_ptr())
    out = torch.empty(b, c, size[0], size[1], dtype=torch.float32, device=x4.device)
    check(lib.gsd_op_image_affine(C.c_void_p(x4.data_ptr()), bptr, base_batch, int(base is not None), b, c, hr, wr,
                                  int(size[0]), int(size[1]), s8, t8, C.c_void_p(out.data_ptr()), int(split_fingers),
                                  x4.device.index or 0, C.c_void_p(torch.cuda.current_stream(x4.device).cuda_stream)),
          "gsd_op_image_affine")
    return out[0] if squeeze else out


def get_difference_image(tactile_image, base_tactile_image):
    """(tactile - base + 255) / 2, base broadcast over the batch (image_utils.py:6-10)."""
    return _affine(tactile_image, tactile_image.shape[-2:], [1.0], [0.0], base=base_tactile_image)


def sample_multi_channel_image_to_desired_size(MC_image, desired_size: Tuple[int, int], interp_method='area'):
    """F.interpolate(size, mode='area') == adaptive average pooling (image_utils.py:12-15)."""
    if interp_method != 'area':
        raise NotImplementedError("only interp_method='area' (the shipped configuration, "
                                  "config_unet_bigdata.py:25) has a kernel; no fallback path")
    return _affine(MC_image, tuple(desired_size), [1.0], [0.0])


def blur_depth_images(depth, depth_image_blur_kernel):
    """torchvision.transforms.functional.gaussian_blur(depth, kernel_size=k) (image_utils.py:17-19): depthwise Gaussian
    with torchvision's default sigma 0.3*((k-1)*0.5-1)+0.8 and reflect padding."""
    if not depth.is_cuda:
        raise RuntimeError("gelslim_depth_b200 processing helpers run on a B200 only (no CPU fallback)")
    k = int(depth_image_blur_kernel)
    d = depth.contiguous().float()
    h, w = d.shape[-2:]
    out = torch.empty_like(d)
    check(lib.gsd_op_gaussian_blur(C.c_void_p(d.data_ptr()), d.numel() // (h * w), h, w, k, 0.0, C.c_void_p(out.data_ptr()),
                                   C.c_void_p(torch.cuda.current_stream(d.device).cuda_stream)), "gsd_op_gaussian_blur")
    return out
