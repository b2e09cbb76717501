"""Generate tests/golden/train_curve.pt: 200 steps of the UNMODIFIED reference U-Net with the trainer's init / Adam
hyper-parameters (train_unet.py:248-250,306) on a small fixed synthetic problem.  Build container only (imports /root/reference)."""
import sys, time, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/reference')
import torch, oracle
from gelslim_depth.models.unet import UNet
torch.set_num_threads(8)
torch.manual_seed(0)
net = UNet(3, 1)
sd = oracle.trainer_init_state_dict(net.state_dict(), seed=7)     # train_unet.py:248-250
net.load_state_dict(sd)
g = torch.Generator().manual_seed(33)
N = 16
X = torch.rand(N, 3, 32, 43, generator=g)
# smooth synthetic targets in the 'min_max_to_0_-1' range (train_unet.py:47): a fixed random linear functional of the input
T = -0.9 * torch.sigmoid(4 * (X.mean(dim=1, keepdim=True) - 0.5) + torch.nn.functional.avg_pool2d(X[:, :1], 5, 1, 2) - 0.5)
opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-6)
net.train()
losses = []
t0 = time.time()
for step in range(200):
    idx = torch.arange(4) + 4 * (step % 4)
    opt.zero_grad()
    out = net(x=X[idx])
    loss = torch.mean((out - T[idx]) ** 2)
    loss.backward()
    opt.step()
    losses.append(float(loss.detach()))
    if step % 20 == 0: print(step, losses[-1], time.time() - t0, flush=True)
torch.save({"X": X, "T": T, "losses": losses, "digest": oracle.state_dict_digest(sd), "module_seed": 0, "init_seed": 7,
            "batch": 4, "steps": 200}, '/root/repo/tests/golden/train_curve.pt')
print("done", losses[::20])
