"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck): tiny geometry, one call each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, oracle
from gelslim_depth_b200.models.unet import UNet
from gelslim_depth_b200.train.engine import FusedTrainer
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = UNet(6, 2)
net.load_state_dict(oracle.conditioned_state_dict(net.state_dict(), 1))
net = net.to(dev).eval()
x = torch.rand(2, 6, 48, 59, device=dev)
y = net(x=x)                                   # halo + tap + head-fused kernels, odd geometry
net.set_precision("fp32"); y32 = net(x=x); net.set_precision("bf16")
net.train()
ft = FusedTrainer(net)
loss = ft.step(x, -0.9 * torch.rand(2, 2, 48, 59, device=dev))   # stats epilogue, bn, wgrad, dgrad, convT bwd, adam
torch.cuda.synchronize()
print("ok", float(y.mean()), float(y32.mean()), float(loss))
