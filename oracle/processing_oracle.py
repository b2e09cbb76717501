"""CPU restatement of the processing utilities on the depth-estimation path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations relative to /root/reference/.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch


def get_difference_image(tactile_image, base_tactile_image):
    """image_utils.py:6-10 -- (img - base + 255) / 2, base broadcast over the batch."""
    return (tactile_image - base_tactile_image + 255.0) / 2.0


def split_fingers(images):
    """general_dataset.py:71 -- (N,6,H,W) -> (2N,3,H,W): all Left fingers first, then all Right."""
    return torch.cat([images[:, 0:3], images[:, 3:6]], dim=0)


def _area_matrix(n_in: int, n_out: int, dtype) -> torch.Tensor:
    """Averaging matrix of adaptive_avg_pool (what F.interpolate(mode='area') dispatches to,
    image_utils.py:14): output bin i covers input [floor(i*in/out), ceil((i+1)*in/out))."""
    m = torch.zeros(n_out, n_in, dtype=dtype)
    for i in range(n_out):
        lo = (i * n_in) // n_out
        hi = -((-(i + 1) * n_in) // n_out)
        m[i, lo:hi] = 1.0 / (hi - lo)
    return m


def area_resample(img: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """sample_multi_channel_image_to_desired_size(..., 'area'), image_utils.py:12-15."""
    ah = _area_matrix(img.shape[-2], size[0], img.dtype)
    aw = _area_matrix(img.shape[-1], size[1], img.dtype)
    return torch.einsum("oh,nchw,pw->ncop", ah, img, aw)


def _image_norm_constants(method: str, norm_scale: float, params):
    """normalization_utils.py:5-22."""
    if "0_255" not in method:
        mins, maxes, means, stds = params
    if method == "min_max_to_-1_1":
        scale = norm_scale
        bias = (0.5 * (torch.tensor(maxes) + torch.tensor(mins))).tolist()
        den = (torch.tensor(maxes) - torch.tensor(mins)).tolist()
    elif method == "mean_std":
        scale, bias, den = 1.0, list(means), list(stds)
    elif method == "0_255_to_-1_1":
        scale, bias, den = 2.0, [127.5], [255.0]
    elif method == "0_255_to_0_1":
        scale, bias, den = 1.0, [0.0], [255.0]
    else:
        raise ValueError(method)
    return scale, bias, den


def normalize_tactile_image(img, method, norm_scale, params=None):
    """normalization_utils.py:4-35: per channel scale*(x - bias[min(i,len-1)])/den[min(i,len-1)]."""
    scale, bias, den = _image_norm_constants(method, norm_scale, params)
    out = torch.zeros_like(img)
    cdim = 0 if img.dim() == 3 else 1
    for i in range(img.shape[cdim]):
        b, d = bias[min(i, len(bias) - 1)], den[min(i, len(den) - 1)]
        if cdim == 0:
            out[i] = scale * (img[i] - b) / d
        else:
            out[:, i] = scale * (img[:, i] - b) / d
    return out


def _depth_norm_constants(method: str, norm_scale: float, params: Sequence[float]):
    """normalization_utils.py:71-96 / 102-127; params may have 2 or 4 entries."""
    p = list(params) if params is not None else []
    mn = p[0] if len(p) > 0 else None
    mx = p[1] if len(p) > 1 else None
    mean = p[2] if len(p) > 2 else None
    std = p[3] if len(p) > 3 else None
    if method == "min_max_to_-1_1":
        return norm_scale, 0.5 * (mx + mn), (mx - mn)
    if method == "mean_std":
        return 1.0, mean, std
    if method == "min_max_to_0_1":
        return norm_scale, mn, mx - mn
    if method == "min_max_to_0_-1":
        return -norm_scale, mn, mx - mn
    raise ValueError(method)


def normalize_depth_image(depth, method, norm_scale, params=None):
    """normalization_utils.py:70-99."""
    scale, bias, den = _depth_norm_constants(method, norm_scale, params)
    return scale * (depth - bias) / den


def denormalize_depth_image(depth, method, norm_scale, params=None):
    """normalization_utils.py:101-130."""
    scale, bias, den = _depth_norm_constants(method, norm_scale, params)
    return (depth * den) / scale + bias


def _cfg(config, *names):
    for n in names:
        if hasattr(config, n):
            return getattr(config, n)
    raise AttributeError(names[0])


def predict_depth_from_RGB(images, model_fn, output_size, config):
    """complete_prediction.py:4-10 (working copy: test_utils/test_depth_estimation.py:14-20).
    ``model_fn`` is any callable image -> normalised depth (e.g. a closure over unet_forward).
    Both spellings of the image-normalisation attributes are accepted."""
    method = _cfg(config, "tactile_normalization_method", "image_normalization_method")
    params = _cfg(config, "tactile_normalization_parameters", "image_normalization_parameters")
    x = area_resample(images, tuple(config.input_tactile_image_size))
    x = normalize_tactile_image(x, method, config.norm_scale, params)
    d = model_fn(x)
    d = denormalize_depth_image(d, config.depth_normalization_method, config.norm_scale,
                                config.depth_normalization_parameters)
    return area_resample(d, tuple(output_size))


def blur_depth_images(depth, kernel_size: int):
    """image_utils.py:17-19 -> torchvision gaussian_blur(kernel_size=k) restated: separable Gaussian with
    sigma = 0.3*((k-1)*0.5-1)+0.8 (torchvision's default), reflect padding, applied per plane with explicit loops over
    the taps (independent of torchvision and of F.conv2d)."""
    k = int(kernel_size)
    sigma = 0.3 * ((k - 1) * 0.5 - 1) + 0.8
    xs = torch.linspace(-(k - 1) * 0.5, (k - 1) * 0.5, k, dtype=torch.float64)
    w = torch.exp(-0.5 * (xs / sigma) ** 2)
    w = (w / w.sum()).to(depth.dtype)
    r = k // 2
    h, wd = depth.shape[-2:]

    def reflect(i, n):
        return -i if i < 0 else (2 * n - 2 - i if i >= n else i)

    rows = torch.zeros_like(depth)
    for t in range(k):                                   # horizontal pass
        idx = torch.tensor([reflect(x + t - r, wd) for x in range(wd)])
        rows = rows + w[t] * depth[..., idx]
    out = torch.zeros_like(depth)
    for t in range(k):                                   # vertical pass
        idx = torch.tensor([reflect(y + t - r, h) for y in range(h)])
        out = out + w[t] * rows[..., idx, :]
    return out


def preprocess_object_tensors(tactile_image, depth_image, base_tactile_image, input_tactile_image_size, image_normalization_method,
                              image_normalization_parameters, depth_normalization_method, depth_normalization_parameters, norm_scale,
                              separate_fingers=True, use_difference_image=True, depth_image_blur_kernel=1):
    """general_dataset.py:61-97 (`load_object_dataset`) followed by :211-215 (`normalize_sample`) on whole tensors."""
    size = tuple(input_tactile_image_size)
    x = get_difference_image(tactile_image, base_tactile_image) if use_difference_image else tactile_image
    d = depth_image
    if separate_fingers:
        x = split_fingers(x)
        d = torch.cat((d[:, 0:1], d[:, 1:2]), dim=0)
    x = area_resample(x, size)
    d = area_resample(d, size)
    if depth_image_blur_kernel > 1:
        d = blur_depth_images(d, depth_image_blur_kernel)
    return {"tactile_image": normalize_tactile_image(x, image_normalization_method, norm_scale, image_normalization_parameters),
            "depth_image": normalize_depth_image(d, depth_normalization_method, norm_scale, depth_normalization_parameters)}
