"""N training steps of the bench workload (UNet(6,2), 6x320x427, batch B) -- the command ncu wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gelslim_depth_b200.models.unet import UNet
from gelslim_depth_b200.train.engine import FusedTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = UNet(6, 2).to(dev).train()
ft = FusedTrainer(net)
x = torch.rand(B, 6, 320, 427, device=dev)
t = -0.9 * torch.rand(B, 2, 320, 427, device=dev)
for _ in range(n):
    loss = ft.step(x, t)
torch.cuda.synchronize()
print("ok", float(loss))
