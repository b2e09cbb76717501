import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
from gelslim_depth_b200.models.unet import UNet
dev = torch.device("cuda:0")
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
for dims, (h, w) in (((64, 128), (40, 53)), ((64, 128), (32, 64)), ((64, 128, 256), (48, 64))):
    torch.manual_seed(3)
    net = UNet(3, 1, layer_dimensions=list(dims))
    sd = oracle.conditioned_state_dict(net.state_dict(), seed=4)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, h, w, generator=g); tgt = -0.9 * torch.rand(2, 1, h, w, generator=g)
    tr = oracle.TrainOracle(sd)
    loss_ref, grads_ref, stats_ref, y_ref = tr.loss_and_grads(x, tgt)
    net = net.to(dev).train()
    y = net(x=x.to(dev))
    loss = torch.mean((y - tgt.to(dev)) ** 2); loss.backward()
    print(dims, h, w, "fwd", rel(y.detach(), y_ref))
    for name, p in net.named_parameters():
        print(f"   {name:50s} {rel(p.grad, grads_ref[name]):.4f}")
