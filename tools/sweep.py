"""BASELINE configs[4]: batch-sharded inference sweep, batch 1..1024 on this GPU (run one process per GPU for N>1),
device-resident frames, plus the pipelined host path at a few chunk sizes.  One JSON line per batch."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gelslim_depth_b200.models.unet import UNet
from gelslim_depth_b200.engine import make_prepost

def run(batches=(1, 2, 4, 8, 16, 32, 64, 128, 256), host_chunks=(4, 8, 16, 32)):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = UNet(6, 2).to(dev).eval()
    H, W = 320, 427
    base = torch.randint(0, 256, (1, 6, H, W), dtype=torch.uint8).float().to(dev)
    pp = make_prepost(6, (H, W), (H, W), use_diff=True, in_scale=[1 / 255.0], out_scale=1.9180814027786255 / -0.9,
                      out_shift=-1.9180814027786255)
    for B in batches:
        x = torch.randint(0, 256, (B, 6, H, W), dtype=torch.uint8).float().to(dev)
        y = torch.empty(B, 2, H, W, device=dev)
        plan = net.plan_for(B, H, W, dev)
        packed = net.packed_weights(plan)
        for _ in range(3):
            plan.forward(x, base, pp, y, packed)
        torch.cuda.synchronize()
        iters = max(5, min(200, int(2000 / B)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            plan.forward(x, base, pp, y, packed)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        rec = {"batch": B, "ms_per_batch": ms, "frames_per_s": B / ms * 1e3, "tensor_frac_of_sustained": B / ms * 1e3 * 200.117 / 1e3 / 1386.1}
        if B == 64:
            xh, yh = x.cpu().pin_memory(), torch.empty(B, 2, H, W).pin_memory()
            for ch in host_chunks:
                plan.set_chunk(ch)
                for _ in range(2):
                    plan.forward_host(xh, base, pp, yh, x, y, packed)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(5):
                    plan.forward_host(xh, base, pp, yh, x, y, packed)
                rec[f"e2e_fps_chunk{ch}"] = B * 5 / (time.perf_counter() - t0)
            plan.set_chunk(B)
        print(json.dumps(rec), flush=True)
        del plan, x, y
        net._plans.clear()
        torch.cuda.empty_cache()

if __name__ == "__main__":
    run()
