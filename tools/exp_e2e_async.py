"""Host pipeline across batches: gsd_forward_host_async on rotating staging slots vs the blocking call.
usage: python tools/exp_e2e_async.py  (B200; prints frames/s per (chunk, slots, input type))"""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
from gelslim_depth_b200.models.unet import UNet
from gelslim_depth_b200.engine import make_prepost
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = UNet(6, 2).to(dev).eval()
H, W, B, STEPS = 320, 427, 64, 10
base = torch.randint(0, 256, (1, 6, H, W), dtype=torch.uint8).float().to(dev)
kw = dict(use_diff=True, in_scale=[1 / 255.0], out_scale=1.9180814027786255 / -0.9, out_shift=-1.9180814027786255)
x8 = torch.randint(0, 256, (B, 6, H, W), dtype=torch.uint8)
plan = net.plan_for(B, H, W, dev)
packed = net.packed_weights(plan)
for u8 in (False, True):
    pp = make_prepost(6, (H, W), (H, W), input_u8=u8, **kw)
    xh = (x8 if u8 else x8.float()).pin_memory()
    for slots in (2, 3):
        yh = [torch.empty(B, 2, H, W).pin_memory() for _ in range(slots)]
        xd = [torch.empty_like(xh, device=dev) for _ in range(slots)]
        yd = [torch.empty(B, 2, H, W, device=dev) for _ in range(slots)]
        for chunk in (64, 32, 16):
            plan.set_chunk(chunk)

            def run():
                for k in range(STEPS):
                    s = k % slots
                    plan.forward_host_async(xh, base, pp, yh[s], xd[s], yd[s], packed, slot=s)
                for s in range(slots):
                    plan.host_wait(s)
            run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run()
            dt = (time.perf_counter() - t0) / STEPS
            print("u8" if u8 else "f32", "slots", slots, "chunk", chunk, round(B / dt), "frames/s", round(dt * 1e3, 2), "ms", flush=True)
        del yh, xd, yd
