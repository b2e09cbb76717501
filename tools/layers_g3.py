"""Per-launch CUDA-event times of the shipped G3 pipeline (UNet(3,1) @160x213, 2 fingers per 320x427 frame pair)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gelslim_depth_b200.models.unet import UNet
from gelslim_depth_b200.engine import make_prepost
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64          # frame pairs
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = UNet(3, 1).to(dev).eval()
x = torch.randint(0, 256, (P, 6, 320, 427), dtype=torch.uint8, device=dev)
base = torch.rand(1, 6, 320, 427, device=dev) * 255
y = torch.empty(2 * P, 1, 320, 427, device=dev)
pp = make_prepost(3, (320, 427), (320, 427), use_diff=True, in_scale=[1 / 255.0], split_fingers=True, input_u8=True)
plan = net.plan_for(2 * P, 160, 213, dev)
packed = net.packed_weights(plan)
for _ in range(5):
    prof = plan.forward_profiled(x, base, pp, y, packed)
names = ["prologue", "inc.0", "inc.3"] + [f"down.{i}.{j}" for i in range(4) for j in (0, 3)] + \
        [f"up.{i}.{n}" for i in range(4) for n in ("up", "conv.0", "conv.3")] + ["head+resample"]
tot = 0
for n, (ms, fl) in zip(names, prof):
    tot += ms
    print(f"{n:14s} {ms*1e3:8.1f} us  {fl/ms/1e9 if fl else 0:8.1f} TFLOP/s")
print("sum", tot * 1e3, "us")
