"""Real-time frame ingest for the depth-estimation hot path (BASELINE configs[2]; README.md:151-171 of the
reference: "in whatever way you acquire tactile images (ROS, webcam, etc.) ...").

`DepthStream` owns a ring of pinned host slots.  `push(frame)` copies one camera frame (uint8 HWC as a camera driver /
cv2 delivers it, uint8 CHW, or float CHW in 0..255) into the next slot and replays that slot's CUDA graph:

    H2D of the frame -> gsd_forward (difference image, Left/Right split, area resampling, normalisation, the 23 U-Net
    layers, depth de-normalisation, all library kernels chained with programmatic dependent launch) -> D2H of the depth map

`result(ticket)` waits for that replay's event and returns the depth map (mm) that now sits in the slot's pinned output
buffer.  One graph replay per frame pair: 25 launches, no allocation, no Python on the device path."""
from __future__ import annotations

import time
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .engine import Plan
from .models.unet import UNet
from .processing_utils.complete_prediction import _prepost_from_config

_LAYOUTS = {"hwc_u8": 2, "chw_u8": 1, "chw_f32": 0}


class DepthStream:
    def __init__(self, model: UNet, config, frame_hw: Tuple[int, int], base_tactile_image: Optional[torch.Tensor] = None,
                 output_size: Optional[Tuple[int, int]] = None, layout: str = "hwc_u8", frame_pairs: bool = False, slots: int = 4,
                 device: Optional[torch.device] = None):
        """model: eval-mode gelslim_depth_b200 UNet; config: the reference's config object (config_unet_bigdata.py);
        frame_hw: camera resolution; base_tactile_image: undeformed reference frame (CHW float or uint8 in the frame
        layout) -> the difference image is computed inside the first kernel; frame_pairs: frames hold Left|Right
        (2 x n_channels channels, general_dataset.py:71) and the result has one depth channel per finger."""
        if not isinstance(model, UNet):
            raise TypeError("DepthStream needs a gelslim_depth_b200.models.unet.UNet (no fallback path)")
        if model.training:
            raise RuntimeError("DepthStream runs the eval-mode network: call model.eval() first")
        if layout not in _LAYOUTS:
            raise ValueError(f"layout must be one of {sorted(_LAYOUTS)}")
        self.model, self.config, self.layout = model, config, layout
        self.device = device or next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("DepthStream runs on a B200 only (no CPU fallback)")
        H, W = int(frame_hw[0]), int(frame_hw[1])
        self.frame_hw = (H, W)
        self.out_hw = tuple(output_size) if output_size is not None else (H, W)
        self.frame_channels = model.n_channels * (2 if frame_pairs else 1)
        self.frame_pairs = bool(frame_pairs)
        self.net_batch = 2 if frame_pairs else 1
        self.net_hw = tuple(config.input_tactile_image_size)
        use_diff = base_tactile_image is not None
        self.pp = _prepost_from_config(config, model.n_channels, (H, W), self.out_hw, use_diff=use_diff, base_batch=1,
                                       split_fingers=frame_pairs, input_u8=_LAYOUTS[layout])
        self.base = self._base_to_device(base_tactile_image) if use_diff else None
        u8 = layout != "chw_f32"
        shape = (1, H, W, self.frame_channels) if layout == "hwc_u8" else (1, self.frame_channels, H, W)
        dt = torch.uint8 if u8 else torch.float32
        # A private plan (workspace) and a private packed-weight buffer: the captured graphs replay on `self.stream`
        # while eager model.run() calls of the same geometry may be in flight on the caller's stream -- sharing the
        # model's PlanCache entry or its packed buffer would make the two race on one workspace.
        dtype = _lib.DTYPE_FP32 if model.precision == "fp32" else _lib.DTYPE_BF16
        self.plan = Plan(self.net_batch, model.n_channels, self.net_hw[0], self.net_hw[1], model.n_classes,
                         model.layer_dimensions, self.device, dtype=dtype)
        self.packed = torch.empty(self.plan.packed_bytes, dtype=torch.uint8, device=self.device)
        self.stream = torch.cuda.Stream(self.device)
        self.refresh_weights()
        self._slots: List[dict] = []
        for _ in range(max(1, int(slots))):
            s = {"x_host": torch.empty(shape, dtype=dt).pin_memory(), "x_dev": torch.empty(shape, dtype=dt, device=self.device),
                 "y_dev": torch.empty(self.net_batch, model.n_classes, *self.out_hw, device=self.device),
                 "y_host": torch.empty(self.net_batch, model.n_classes, *self.out_hw).pin_memory(),
                 "done": torch.cuda.Event(), "graph": None, "t_push": 0.0}
            self._slots.append(s)
        with torch.cuda.stream(self.stream):
            for s in self._slots[:1]:
                for _ in range(2):                          # warm-up outside capture (lazy module / attribute set-up)
                    self._enqueue(s)
        self.stream.synchronize()
        for s in self._slots:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.stream):
                self._enqueue(s)
            s["graph"] = g
        self._next = 0
        self.latencies_ms: List[float] = []

    # ------------------------------------------------------------------ helpers
    def _base_to_device(self, base: torch.Tensor) -> torch.Tensor:
        b = torch.as_tensor(base)
        if b.dim() == 4:
            b = b[0]
        if b.shape[-1] == self.frame_channels and b.shape[0] != self.frame_channels:    # HWC -> CHW
            b = b.permute(2, 0, 1)
        if tuple(b.shape) != (self.frame_channels, *self.frame_hw):
            raise ValueError(f"base image must be ({self.frame_channels}, {self.frame_hw[0]}, {self.frame_hw[1]}), got {tuple(b.shape)}")
        return b.to(self.device, torch.float32).contiguous()[None]

    def _enqueue(self, s: dict):
        s["x_dev"].copy_(s["x_host"], non_blocking=True)
        self.plan.forward(s["x_dev"], self.base, self.pp, s["y_dev"], self.packed)
        s["y_host"].copy_(s["y_dev"], non_blocking=True)

    def refresh_weights(self):
        """Call after the model's parameters changed (load_state_dict, EMA swap): re-packs, on the stream the captured
        graphs replay on, into the buffer they read."""
        self.stream.wait_stream(torch.cuda.current_stream(self.device))      # the caller's parameter writes come first
        with torch.cuda.stream(self.stream):
            self.plan.pack([p.detach() for p in self.model.parameters()], self.model._bn_buffers(), self.packed)
        self.stream.synchronize()

    # ------------------------------------------------------------------ streaming API
    def push(self, frame) -> int:
        """Copy `frame` into the next ring slot and launch its graph.  Returns a ticket for result()."""
        i = self._next
        s = self._slots[i]
        self._next = (i + 1) % len(self._slots)
        s["done"].synchronize()                              # slot still in flight from a previous lap?
        f = torch.as_tensor(frame)
        if f.dim() == 3:
            f = f[None]
        if f.shape != s["x_host"].shape or f.dtype != s["x_host"].dtype:
            raise ValueError(f"frame must be {tuple(s['x_host'].shape)} {s['x_host'].dtype} for layout '{self.layout}', "
                             f"got {tuple(f.shape)} {f.dtype}")
        s["t_push"] = time.perf_counter()
        s["x_host"].copy_(f)
        with torch.cuda.stream(self.stream):
            s["graph"].replay()
            s["done"].record(self.stream)
        return i

    def acquire(self):
        """Zero-copy ingest: -> (ticket, pinned input buffer of the next ring slot).  A camera driver writes the frame
        straight into the buffer (shape / dtype of the configured layout), then calls submit(ticket)."""
        i = self._next
        s = self._slots[i]
        self._next = (i + 1) % len(self._slots)
        s["done"].synchronize()
        return i, s["x_host"][0]

    def submit(self, ticket: int) -> int:
        """Launch the graph of a slot whose input buffer was filled in place (see acquire())."""
        s = self._slots[ticket]
        s["t_push"] = time.perf_counter()
        with torch.cuda.stream(self.stream):
            s["graph"].replay()
            s["done"].record(self.stream)
        return ticket

    def result(self, ticket: int) -> torch.Tensor:
        """Depth map(s) in mm for the frame pushed with `ticket`: (n_classes, H, W) for single-finger frames,
        (2, H, W) [Left, Right] for frame pairs.  A view of the slot's pinned buffer, valid until the slot is reused."""
        s = self._slots[ticket]
        s["done"].synchronize()
        self.latencies_ms.append((time.perf_counter() - s["t_push"]) * 1e3)
        y = s["y_host"]
        return y[:, 0] if self.frame_pairs else y[0]

    def __call__(self, frame) -> torch.Tensor:
        return self.result(self.push(frame))

    def latency_percentiles(self, qs: Sequence[float] = (0.5, 0.99)) -> List[float]:
        lat = sorted(self.latencies_ms)
        if not lat:
            return [float("nan")] * len(qs)
        return [lat[min(len(lat) - 1, int(q * len(lat)))] for q in qs]
