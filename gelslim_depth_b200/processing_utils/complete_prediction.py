"""Drop-in for gelslim_depth/processing_utils/complete_prediction.py:4-10.

predict_depth_from_RGB(images, model, output_size, config): area-resample -> normalise -> U-Net ->
de-normalise -> area-resample.  With a gelslim_depth_b200 UNet the four processing steps are folded
into the network's first and last kernels (one gsd_forward call); nothing is materialised between
them.  `predict_depth_from_frames` additionally fuses the caller-side difference image and the
Left/Right channel split (README.md:155-170, general_dataset.py:71)."""
from __future__ import annotations

import torch

from ..engine import make_prepost
from ..models.unet import UNet
from .normalization_utils import depth_affine_constants, image_affine_constants


def _cfg(config, *names):
    for n in names:
        if hasattr(config, n):
            return getattr(config, n)
    raise AttributeError(f"config has none of {names}")


def _prepost_from_config(config, n_channels, raw_hw, output_size, use_diff=False, base_batch=1, split_fingers=False,
                         input_u8=False):
    # complete_prediction.py:6 reads `tactile_normalization_*`; the shipped/generated configs define
    # `image_normalization_*` (config_unet_bigdata.py:39-40, test_depth_estimation.py:16): accept both.
    method = _cfg(config, "tactile_normalization_method", "image_normalization_method")
    params = _cfg(config, "tactile_normalization_parameters", "image_normalization_parameters")
    if getattr(config, "interp_method", "area") != "area":
        raise NotImplementedError("only interp_method='area' has a kernel (no fallback path)")
    in_scale, in_shift = image_affine_constants(method, config.norm_scale, params)
    scale, bias, den = depth_affine_constants(config.depth_normalization_method, config.norm_scale,
                                              config.depth_normalization_parameters)
    return make_prepost(n_channels, raw_hw, tuple(output_size), use_diff=use_diff, base_batch=base_batch,
                        in_scale=in_scale, in_shift=in_shift, out_scale=den / scale, out_shift=bias,
                        split_fingers=split_fingers, input_u8=input_u8)


def predict_depth_from_RGB(images, model, output_size, config):
    if not isinstance(model, UNet):
        raise TypeError("predict_depth_from_RGB needs a gelslim_depth_b200.models.unet.UNet (no fallback path)")
    pp = _prepost_from_config(config, model.n_channels, tuple(images.shape[-2:]), output_size)
    return model.run(images, pp=pp, net_hw=tuple(config.input_tactile_image_size))


def predict_depth_from_frames(tactile_frames, base_tactile_image, model, output_size, config):
    """Raw camera frames (N, n_channels, H, W) in 0..255 + the undeformed base image -> depth in mm,
    with get_difference_image fused into the prologue as well."""
    if not isinstance(model, UNet):
        raise TypeError("predict_depth_from_frames needs a gelslim_depth_b200.models.unet.UNet")
    base = base_tactile_image if base_tactile_image.dim() == 4 else base_tactile_image[None]
    pp = _prepost_from_config(config, model.n_channels, tuple(tactile_frames.shape[-2:]), output_size,
                              use_diff=True, base_batch=base.shape[0], input_u8=tactile_frames.dtype == torch.uint8)
    return model.run(tactile_frames, pp=pp, base=base.to(tactile_frames.device),
                     net_hw=tuple(config.input_tactile_image_size))


def predict_depth_from_frame_pairs(tactile_frames, base_tactile_image, model, output_size, config):
    """The shipped pipeline end to end for Left|Right frame pairs (README.md:155-171, general_dataset.py:71):
    tactile_frames (N, 6, H, W) float or uint8 camera frames, base (1|N, 6, H, W) -> depth (N, 2, H, W) in mm with
    channel 0 = Left finger, channel 1 = Right finger.  `model` is the 3-channel / 1-class U-Net of
    train_unet.py:235; the Left/Right split, difference image, resampling and normalisation all happen inside the
    first kernel, the two fingers run as one batch of 2N."""
    if not isinstance(model, UNet):
        raise TypeError("predict_depth_from_frame_pairs needs a gelslim_depth_b200.models.unet.UNet")
    n = tactile_frames.shape[0]
    base = base_tactile_image if base_tactile_image.dim() == 4 else base_tactile_image[None]
    pp = _prepost_from_config(config, model.n_channels, tuple(tactile_frames.shape[-2:]), output_size, use_diff=True,
                              base_batch=base.shape[0], split_fingers=True, input_u8=tactile_frames.dtype == torch.uint8)
    y = model.run(tactile_frames, pp=pp, base=base.to(tactile_frames.device).float(), net_hw=tuple(config.input_tactile_image_size))
    # (2N, 1, H, W) [all Left, then all Right] -> (N, 2, H, W): a strided view, no copy
    return y.view(2, n, *y.shape[2:]).permute(1, 0, 2, 3)
