"""Is the gap between the sum of stand-alone kernel times and the back-to-back step time the power cap?
Times the batch-64 forward in loops of 1..100 steps, and single steps separated by idle gaps."""
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
x = torch.randint(0, 256, (B, 6, H, W), dtype=torch.uint8).float().to(dev)
y = torch.empty(B, 2, H, W, device=dev)
plan = net.plan_for(B, H, W, dev)
packed = net.packed_weights(plan)
for _ in range(3):
    plan.forward(x, base, pp, y, packed)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for gap in (0.5, 0.05):
    ts = []
    for _ in range(6):
        time.sleep(gap)
        e0.record(); plan.forward(x, base, pp, y, packed); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 3))
    print("single steps after", gap, "s idle:", ts, flush=True)
for n in (1, 2, 3, 5, 10, 30, 100, 300):
    time.sleep(0.5)
    e0.record()
    for _ in range(n):
        plan.forward(x, base, pp, y, packed)
    e1.record(); torch.cuda.synchronize()
    print("loop", n, "steps:", round(e0.elapsed_time(e1) / n, 3), "ms/step", flush=True)
prof = plan.forward_profiled(x, base, pp, y, packed)
print("profiled sum", round(sum(m for m, f in prof), 3))
