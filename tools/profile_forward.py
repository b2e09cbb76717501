"""Two forwards of the bench workload (UNet(6,2), 6x320x427, batch B) -- the command ncu wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gelslim_depth_b200.models.unet import UNet
from gelslim_depth_b200.engine import make_prepost

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = UNet(6, 2).to(dev).eval()
x = torch.randint(0, 256, (B, 6, 320, 427), dtype=torch.uint8).float().to(dev)
base = torch.randint(0, 256, (1, 6, 320, 427), dtype=torch.uint8).float().to(dev)
y = torch.empty(B, 2, 320, 427, device=dev)
pp = make_prepost(6, (320, 427), (320, 427), use_diff=True, in_scale=[1 / 255.0], out_scale=1.9180814027786255 / -0.9,
                  out_shift=-1.9180814027786255)
plan = net.plan_for(B, 320, 427, dev)
packed = net.packed_weights(plan)
for _ in range(n):
    plan.forward(x, base, pp, y, packed)
torch.cuda.synchronize()
print("ok", float(y.mean()))
