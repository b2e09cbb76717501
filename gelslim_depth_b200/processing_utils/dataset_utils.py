"""GPU-side form of the preprocessing that gelslim_depth/datasets/general_dataset.py applies to every object file
(`load_object_dataset`, :61-97) and to every sample (`normalize_sample`, :211-215), i.e. the step immediately before the
training hot path.  The reference does it with ~6 ATen ops per tensor on the CPU at load time plus a per-sample
normalisation in `__getitem__`; here the Left/Right split, difference image, area down-sampling and normalisation of a
whole object tensor are ONE launch of the library's image_affine kernel per tensor (plus one blur launch when
`depth_image_blur_kernel > 1`).  Only the arithmetic is mirrored: file discovery, `torch.load`, shuffling and the
train/validation split stay with the caller (out of scope)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .image_utils import _affine, blur_depth_images
from .normalization_utils import depth_affine_constants, image_affine_constants


def preprocess_object_tensors(tactile_image: torch.Tensor, depth_image: Optional[torch.Tensor],
                              base_tactile_image: Optional[torch.Tensor], input_tactile_image_size: Tuple[int, int],
                              image_normalization_method: str, image_normalization_parameters,
                              depth_normalization_method: str, depth_normalization_parameters, norm_scale: float,
                              separate_fingers: bool = True, use_difference_image: bool = True, interp_method: str = "area",
                              depth_image_blur_kernel: int = 1):
    """tactile_image (N, 6|3, H, W) float 0..255, depth_image (N, 2|1, H, W) mm, base (1|N, 6|3, H, W)
    -> {'tactile_image': (2N|N, 3, h, w) normalised network input, 'depth_image': (2N|N, 1, h, w) normalised target}
    exactly as `GeneralDataset.load_object_dataset` followed by `normalize_sample` on every sample produces them."""
    if interp_method != "area":
        raise NotImplementedError("only interp_method='area' (config_unet_bigdata.py:25) has a kernel; no fallback path")
    if use_difference_image and base_tactile_image is None:
        raise ValueError("use_difference_image=True needs base_tactile_image (general_dataset.py:71)")
    size = tuple(int(v) for v in input_tactile_image_size)
    in_scale, in_shift = image_affine_constants(image_normalization_method, norm_scale, image_normalization_parameters)
    out = {"tactile_image": _affine(tactile_image, size, in_scale, in_shift,
                                    base=base_tactile_image if use_difference_image else None, split_fingers=separate_fingers)}
    if depth_image is not None:
        scale, bias, den = depth_affine_constants(depth_normalization_method, norm_scale, depth_normalization_parameters)
        if depth_image_blur_kernel > 1:
            # the reference blurs the resampled depth in mm and normalises per sample afterwards (:76-79, :214)
            d = blur_depth_images(_affine(depth_image, size, [1.0], [0.0], split_fingers=separate_fingers), depth_image_blur_kernel)
            out["depth_image"] = _affine(d, size, [scale / den], [-scale * bias / den])
        else:
            out["depth_image"] = _affine(depth_image, size, [scale / den], [-scale * bias / den], split_fingers=separate_fingers)
    return out
