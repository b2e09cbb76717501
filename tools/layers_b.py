"""Per-launch CUDA-event times of one forward at batch B (gsd_forward_profiled)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gelslim_depth_b200.models.unet import UNet
from gelslim_depth_b200.engine import make_prepost
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = UNet(6, 2).to(dev).eval()
x = torch.rand(B, 6, 320, 427, device=dev) * 255
base = torch.rand(1, 6, 320, 427, device=dev) * 255
y = torch.empty(B, 2, 320, 427, device=dev)
pp = make_prepost(6, (320, 427), (320, 427), use_diff=True, in_scale=[1 / 255.0])
plan = net.plan_for(B, 320, 427, dev)
packed = net.packed_weights(plan)
for _ in range(5):
    prof = plan.forward_profiled(x, base, pp, y, packed)
fused = len(prof) == 23      # 22 conv launches (prologue fused into the first) + the tail entry
names = (["inc.0+prologue"] if fused else ["prologue", "inc.0"]) + ["inc.3"] + [f"down.{i}.{j}" for i in range(4) for j in (0, 3)] + \
        [f"up.{i}.{n}" for i in range(4) for n in ("up", "conv.0", "conv.3")] + ["tail"]
tot = 0
for n, (ms, fl) in zip(names, prof):
    tot += ms
    print(f"{n:12s} {ms*1e3:8.1f} us  {fl/ms/1e9 if fl else 0:8.1f} TFLOP/s")
print("sum", tot * 1e3, "us")
