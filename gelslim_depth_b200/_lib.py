"""ctypes binding of libgsd_b200.so (C ABI declared in include/gsd_b200.h).

The library is the product: if it is missing or fails to load, importing this module raises --
nothing in this package falls back to PyTorch or CPU arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSD_B200_LIB") or os.path.join(HERE, "libgsd_b200.so")     # override: A/B experiments against an older build

GSD_MAX_DIMS = 8
DTYPE_BF16, DTYPE_FP32 = 0, 1
MODE_INFER, MODE_TRAIN = 0, 1


class Geometry(C.Structure):
    _fields_ = [("batch", C.c_int32), ("in_channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("n_classes", C.c_int32), ("n_dims", C.c_int32), ("dims", C.c_int32 * GSD_MAX_DIMS),
                ("dtype", C.c_int32), ("mode", C.c_int32)]


class PackItem(C.Structure):            # gsd_pack_item
    _fields_ = [("w", C.c_void_p), ("out", C.c_void_p), ("out_dgrad", C.c_void_p), ("mode", C.c_int32), ("O", C.c_int32), ("I", C.c_int32),
                ("Ipad", C.c_int32), ("start", C.c_int64)]


class PrePost(C.Structure):
    _fields_ = [("use_diff", C.c_int32), ("base_batch", C.c_int32), ("raw_height", C.c_int32), ("raw_width", C.c_int32),
                ("out_height", C.c_int32), ("out_width", C.c_int32), ("in_scale", C.c_float * 8),
                ("in_shift", C.c_float * 8), ("out_scale", C.c_float), ("out_shift", C.c_float),
                ("split_fingers", C.c_int32), ("input_u8", C.c_int32)]


class Adam(C.Structure):               # gsd_adam
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("ema_decay", C.c_float), ("grad_scale", C.c_float)]


class OptimizerState(C.Structure):     # gsd_optimizer_state
    _fields_ = [("params", C.c_void_p), ("grads", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("ema", C.c_void_p),
                ("n", C.c_longlong), ("counter", C.c_void_p), ("hp", Adam)]


# gsd_bucket_cb(user, bucket, lo, hi, main_stream, side_stream)
BUCKET_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p)
NULL_CB = C.cast(None, BUCKET_CB)      # "no callback"

# name -> (restype, argtypes): every symbol include/gsd_b200.h declares
SYMBOLS = {
    "gsd_abi_version": (C.c_int, []),
    "gsd_last_error": (C.c_char_p, []),
    "gsd_device_count": (C.c_int, []),
    "gsd_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Geometry), C.c_int]),
    "gsd_plan_destroy": (None, [C.c_void_p]),
    "gsd_plan_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "gsd_plan_packed_bytes": (C.c_size_t, [C.c_void_p]),
    "gsd_plan_num_params": (C.c_int, [C.c_void_p]),
    "gsd_plan_num_bn_buffers": (C.c_int, [C.c_void_p]),
    "gsd_plan_forward_launches": (C.c_int, [C.c_void_p]),
    "gsd_plan_set_chunk": (C.c_int, [C.c_void_p, C.c_int]),
    "gsd_plan_set_chunk_ramp": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "gsd_plan_conv_flops": (C.c_double, [C.c_void_p]),
    "gsd_plan_first_fused": (C.c_int, [C.c_void_p]),
    "gsd_pack_weights": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "gsd_pack_weights_if_changed": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p,
                                              C.c_void_p]),
    "gsd_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PrePost), C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "gsd_forward_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PrePost), C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gsd_forward_host_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PrePost), C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "gsd_forward_host_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "gsd_debug_num_activations": (C.c_int, [C.c_void_p]),
    "gsd_debug_activation_shape": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gsd_debug_read_activation": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gsd_debug_plan_conv3x3": (C.c_int, [C.c_int] * 7 + [C.POINTER(C.c_int)]),
    "gsd_debug_chunk_schedule": (C.c_int, [C.c_int] * 4 + [C.POINTER(C.c_int), C.c_int]),
    "gsd_forward_profiled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PrePost), C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_int,
                                       C.POINTER(C.c_int)]),
    "gsd_op_conv_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p]),
    "gsd_op_conv3x3_halo_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "gsd_op_wgrad3x3_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "gsd_op_conv_auto_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gsd_op_convt_dgrad_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gsd_op_convt_wgrad_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "gsd_op_prologue_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_void_p]),
    "gsd_op_bn_finalize": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gsd_op_negate_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "gsd_op_bn_relu_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gsd_op_mse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gsd_op_head_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gsd_op_head_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gsd_op_bn_bwd_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "gsd_op_bn_bwd_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "gsd_op_adam_ema_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_void_p]),
    "gsd_op_maxpool_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gsd_op_pack_weight": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "gsd_op_unpack_wgrad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "gsd_op_bn_relu_head_fwd": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p, C.c_void_p]),
    "gsd_op_head_bn_bwd": (C.c_int, [C.c_void_p] * 8 + [C.c_double] + [C.c_int] * 4 + [C.c_void_p] * 5),
    "gsd_op_pack_weights_batched": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]),
    "gsd_pack_item_units": (C.c_longlong, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "gsd_op_adam_ema": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_longlong, C.c_float, C.c_longlong, C.c_float, C.c_void_p]),
    "gsd_train_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Geometry), C.c_int]),
    "gsd_debug_train_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Geometry)]),
    "gsd_train_plan_destroy": (None, [C.c_void_p]),
    "gsd_train_plan_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "gsd_train_plan_num_params": (C.c_int, [C.c_void_p]),
    "gsd_train_plan_num_bn": (C.c_int, [C.c_void_p]),
    "gsd_train_plan_param_numel": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.c_int]),
    "gsd_train_plan_launches": (C.c_int, [C.c_void_p]),
    "gsd_train_plan_bind": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p), C.c_void_p]),
    "gsd_train_plan_set_buckets": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "gsd_train_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gsd_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, BUCKET_CB, C.c_void_p]),
    "gsd_adam_ema_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.POINTER(Adam), C.c_void_p,
                                    C.c_void_p]),
    "gsd_train_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(OptimizerState), C.c_void_p, BUCKET_CB,
                                 C.c_void_p]),
    "gsd_op_image_affine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p,
                                      C.c_int, C.c_int, C.c_void_p]),
    "gsd_op_gaussian_blur": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m gelslim_depth_b200.build` (needs nvcc). "
        "gelslim_depth_b200 has no CPU / PyTorch fallback.")

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SYMBOLS.items():
    _fn = getattr(lib, _name)          # AttributeError here == ABI mismatch, fail loudly
    _fn.restype = _res
    _fn.argtypes = _args

if lib.gsd_abi_version() != 1:
    raise ImportError("libgsd_b200.so ABI version mismatch")


class GsdError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise GsdError(f"{what}: {lib.gsd_last_error().decode(errors='replace')} (rc={rc})")
