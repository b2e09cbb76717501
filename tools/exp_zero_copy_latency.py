"""Batch-1 latency experiment: the depth map written by the last kernel STRAIGHT into pinned host memory (no D2H copy
node) and / or the camera frame read by the first conv's producer warps straight from pinned host memory (no H2D node).
Pinned cudaHostAlloc memory is device-accessible under UVA.  Prints p50 / p99 wall-clock per frame pair per variant."""
import json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gelslim_depth_b200.engine import Plan, make_prepost
from gelslim_depth_b200.models.unet import UNet

H, W = 320, 427
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = UNet(6, 2).to(dev).eval()
plan = Plan(1, 6, H, W, 2, net.layer_dimensions, dev)
packed = torch.empty(plan.packed_bytes, dtype=torch.uint8, device=dev)
plan.pack([p.detach() for p in net.parameters()], net._bn_buffers(), packed)
base = torch.randint(0, 256, (1, 6, H, W), dtype=torch.uint8).float().to(dev)
res = {}
for layout, kind in (("hwc_u8", 2), ("chw_u8", 1)):
    shape = (1, H, W, 6) if kind == 2 else (1, 6, H, W)
    xh = torch.randint(0, 256, shape, dtype=torch.uint8).pin_memory()
    xd = torch.empty(shape, dtype=torch.uint8, device=dev)
    yh = torch.empty(1, 2, H, W).pin_memory()
    yd = torch.empty(1, 2, H, W, device=dev)
    pp = make_prepost(6, (H, W), (H, W), use_diff=True, in_scale=[1 / 255.0], out_scale=-2.1312, out_shift=-1.918, input_u8=kind)
    ref = None
    for zc_in in (False, True):
        for zc_out in (False, True):
            s = torch.cuda.Stream()

            def enqueue():
                if not zc_in:
                    xd.copy_(xh, non_blocking=True)
                plan.forward(xh if zc_in else xd, base, pp, yh if zc_out else yd, packed)
                if not zc_out:
                    yh.copy_(yd, non_blocking=True)
            with torch.cuda.stream(s):
                for _ in range(3):
                    enqueue()
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                enqueue()
            for _ in range(20):
                g.replay()
            torch.cuda.synchronize()
            out = yh.clone()
            if ref is None:
                ref = out
            lat = []
            for _ in range(600):
                t0 = time.perf_counter()
                g.replay()
                torch.cuda.synchronize()
                lat.append((time.perf_counter() - t0) * 1e3)
            lat.sort()
            res[f"{layout} zc_in={int(zc_in)} zc_out={int(zc_out)}"] = (round(statistics.median(lat), 4), round(lat[int(0.99 * len(lat)) - 1], 4),
                                                                        bool(torch.equal(out, ref)))
for k, v in res.items():
    print(k, "p50 %.4f p99 %.4f bit-identical %s" % v)
