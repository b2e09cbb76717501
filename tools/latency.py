"""BASELINE configs[2]: real-time streaming inference, batch 1, CUDA-graph replay of the whole hot path
(H2D of one 6x320x427 fp32 frame pair -> gsd_forward with the difference image / normalisation / de-normalisation
fused -> D2H of the 2x320x427 depth map).  Prints p50 / p99 latency per frame pair as one JSON line."""
import json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gelslim_depth_b200.models.unet import UNet
from gelslim_depth_b200.engine import make_prepost

def main(B=1, replays=1000, cin=6, ncls=2, H=320, W=427, net_hw=None):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = UNet(cin, ncls).to(dev).eval()
    nh, nw = net_hw or (H, W)
    xh = torch.randint(0, 256, (B, cin, H, W), dtype=torch.uint8).float().pin_memory()
    base = torch.randint(0, 256, (1, cin, H, W), dtype=torch.uint8).float().to(dev)
    yh = torch.empty(B, ncls, H, W).pin_memory()
    xd, yd = torch.empty_like(xh, device=dev), torch.empty(B, ncls, H, W, device=dev)
    pp = make_prepost(cin, (H, W), (H, W), use_diff=True, in_scale=[1 / 255.0], out_scale=1.9180814027786255 / -0.9,
                      out_shift=-1.9180814027786255)
    plan = net.plan_for(B, nh, nw, dev)
    packed = net.packed_weights(plan)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            xd.copy_(xh, non_blocking=True); plan.forward(xd, base, pp, yd, packed); yh.copy_(yd, non_blocking=True)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        xd.copy_(xh, non_blocking=True)
        plan.forward(xd, base, pp, yd, packed)
        yh.copy_(yd, non_blocking=True)
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    lat = []
    for _ in range(replays):
        t0 = time.perf_counter()
        g.replay()
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    # kernel-only (no copies, no host sync per replay): back-to-back replays timed with events
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2, stream=s):
        plan.forward(xd, base, pp, yd, packed)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(200):
        g2.replay()
    e1.record()
    torch.cuda.synchronize()
    lat.sort()
    out = {"metric": "b1_latency_ms_per_frame_pair", "batch": B, "geometry": f"UNet({cin},{ncls}) {nh}x{nw} (frames {H}x{W})",
           "p50_ms": statistics.median(lat), "p99_ms": lat[int(0.99 * len(lat)) - 1], "min_ms": lat[0], "replays": replays,
           "includes": "H2D frame + fused forward + D2H depth, one CUDA-graph replay + host sync per frame pair",
           "device_only_ms": e0.elapsed_time(e1) / 200, "launches_per_replay": plan.launches}
    print(json.dumps(out), flush=True)
    return out

if __name__ == "__main__":
    main(B=int(sys.argv[1]) if len(sys.argv) > 1 else 1)
    if len(sys.argv) > 2 and sys.argv[2] == "g3":
        main(B=2, cin=3, ncls=1, net_hw=None, H=160, W=213)
