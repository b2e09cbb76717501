import sys, time
sys.path.insert(0, "/root/repo")
import torch
from gelslim_depth_b200.models.unet import UNet
from gelslim_depth_b200.engine import make_prepost
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = UNet(6, 2).to(dev).eval()
H, W, B = 320, 427, 64
base = torch.randint(0, 256, (1, 6, H, W), dtype=torch.uint8).float().to(dev)
pp = make_prepost(6, (H, W), (H, W), use_diff=True, in_scale=[1 / 255.0], out_scale=1.9180814027786255 / -0.9, out_shift=-1.9180814027786255)
x = torch.randint(0, 256, (B, 6, H, W), dtype=torch.uint8).float()
xh, yh = x.pin_memory(), torch.empty(B, 2, H, W).pin_memory()
xd, yd = x.to(dev), torch.empty(B, 2, H, W, device=dev)
plan = net.plan_for(B, H, W, dev)
packed = net.packed_weights(plan)
for (c, f, l) in [(16, 8, 8), (16, 4, 4), (32, 8, 8), (24, 8, 8), (20, 4, 4), (16, 8, 4), (32, 4, 4), (64, 0, 0), (48, 8, 8)]:
    plan.set_chunk(c, first=f, last=l)
    for _ in range(2):
        plan.forward_host(xh, base, pp, yh, xd, yd, packed)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        plan.forward_host(xh, base, pp, yh, xd, yd, packed)
    dt = (time.perf_counter() - t0) / 10
    print(c, f, l, round(B / dt), round(dt * 1e3, 2))
