"""Drop-in for gelslim_depth/processing_utils/normalization_utils.py.

The (scale, bias, denominator) tables restate normalization_utils.py:5-22 and :71-96; the arithmetic
itself is a per-channel affine map executed by the library (image_affine kernel), or folded into
the network prologue / epilogue when called through predict_depth_from_RGB."""
from __future__ import annotations

from typing import Sequence

from .image_utils import _affine


def image_affine_constants(method: str, norm_scale: float, params=None):
    """-> (in_scale[c], in_shift[c]) with x_norm = in_scale*x + in_shift (normalization_utils.py:5-34)."""
    if '0_255' not in method:
        mins, maxes, means, stds = params
    if method == 'min_max_to_-1_1':
        # normalization_utils.py:9 computes `0.5*(tensor).tolist()` == float * list: the reference
        # raises TypeError for this method; same behaviour here rather than guessing the intent.
        raise TypeError("can't multiply sequence by non-int of type 'float' "
                        "(reference normalization_utils.py:9; 'min_max_to_-1_1' is unusable for images)")
    elif method == 'mean_std':
        scale, bias, den = 1.0, list(means), list(stds)
    elif method == '0_255_to_-1_1':
        scale, bias, den = 2.0, [127.5], [255.0]
    elif method == '0_255_to_0_1':
        scale, bias, den = 1.0, [0.0], [255.0]
    else:
        raise UnboundLocalError(f"unknown image_normalization_method {method!r}")   # reference: unbound `scale`
    n = max(len(bias), len(den))
    pick = lambda v, i: v[min(i, len(v) - 1)]   # noqa: E731  (normalization_utils.py:28,34 index rule)
    in_scale = [scale / pick(den, i) for i in range(n)]
    in_shift = [-scale * pick(bias, i) / pick(den, i) for i in range(n)]
    return in_scale, in_shift


def depth_affine_constants(method: str, norm_scale: float, params: Sequence[float] = None):
    """-> (scale, bias, denominator) of normalization_utils.py:71-96 / 102-127 (params of length 2 or 4)."""
    p = list(params) if params is not None else []
    mn = p[0] if len(p) > 0 else None
    mx = p[1] if len(p) > 1 else None
    mean = p[2] if len(p) > 2 else None
    std = p[3] if len(p) > 3 else None
    if method == 'min_max_to_-1_1':
        return norm_scale, 0.5 * (mx + mn), (mx - mn)
    if method == 'mean_std':
        return 1.0, mean, std
    if method == 'min_max_to_0_1':
        return norm_scale, mn, mx - mn
    if method == 'min_max_to_0_-1':
        return -norm_scale, mn, mx - mn
    raise UnboundLocalError(f"unknown depth_normalization_method {method!r}")


def normalize_tactile_image(tactile_image, image_normalization_method, norm_scale, image_normalization_params=None):
    s, t = image_affine_constants(image_normalization_method, norm_scale, image_normalization_params)
    return _affine(tactile_image, tactile_image.shape[-2:], s, t)


def normalize_depth_image(depth_image, depth_normalization_method, norm_scale, depth_normalization_params=None):
    scale, bias, den = depth_affine_constants(depth_normalization_method, norm_scale, depth_normalization_params)
    return _affine(depth_image, depth_image.shape[-2:], [scale / den], [-scale * bias / den])


def denormalize_depth_image(depth_image, depth_normalization_method, norm_scale, depth_normalization_params=None):
    scale, bias, den = depth_affine_constants(depth_normalization_method, norm_scale, depth_normalization_params)
    return _affine(depth_image, depth_image.shape[-2:], [den / scale], [bias])
