import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import oracle
from gelslim_depth_b200.models.unet import UNet
from gelslim_depth_b200.train import engine, ops

dev = torch.device("cuda:0")
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
def nchw(t): return t.permute(0, 3, 1, 2).float().cpu()

torch.manual_seed(3)
net = UNet(3, 1)
sd = oracle.conditioned_state_dict(net.state_dict(), seed=4)
net.load_state_dict(sd)
g = torch.Generator().manual_seed(5)
B, H, W = 2, 64, 85
x = torch.rand(B, 3, H, W, generator=g)
tgt = -0.9 * torch.rand(B, 1, H, W, generator=g)
y_ref, taps, stats = oracle.unet_forward_with_taps(sd, x, training=True)
net = net.to(dev).train()
pw = engine.PackedTrainWeights(net)
y, ctx = engine.train_forward(net, x.to(dev), pw)
print("y", rel(y, y_ref))
names = ["inc"] + [f"down.{i}.maxpool_conv.1" for i in range(4)]
for l, (u1, u2) in enumerate(ctx["enc"]):
    p = names[l]
    print(p, "z1", rel(nchw(u1.z), taps[p + ".double_conv.0"]), "a1", rel(nchw(u1.a), taps[p + ".double_conv.2"]),
          "z2", rel(nchw(u2.z), taps[p + ".double_conv.3"]), "a2", rel(nchw(u2.a), taps[p + ".double_conv.5"]))
for i, (up, y_prev, u, off, u1, u2) in enumerate(ctx["dec"]):
    p = f"up.{i}"
    print(p, "u", rel(nchw(u), taps[p + ".up"]), "z1", rel(nchw(u1.z), taps[p + ".conv.double_conv.0"]),
          "a2", rel(nchw(u2.a), taps[p + ".conv.double_conv.5"]))

# ---- unit checks of the backward ops against torch autograd on identical (bf16-rounded) inputs
def bf(t): return t.to(torch.bfloat16).float()
C = 128
a_in = bf(torch.randn(2, C, 13, 17, generator=g))
zt = a_in.clone().requires_grad_(True)
gamma = (0.5 + torch.rand(C, generator=g)); beta = torch.randn(C, generator=g) * 0.1
bn = torch.nn.BatchNorm2d(C); bn.weight.data = gamma.clone(); bn.bias.data = beta.clone(); bn.train()
a_ref = torch.relu(bn(zt))
da = bf(torch.randn(2, C, 13, 17, generator=g))
a_ref.backward(da)
mean = a_in.mean(dim=(0, 2, 3)); var = a_in.var(dim=(0, 2, 3), unbiased=False); rstd = torch.rsqrt(var + 1e-5)
z_d = a_in.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)
scale = (gamma * rstd).to(dev); shift = (beta - mean * gamma * rstd).to(dev)
a_d, _ = ops.bn_relu_apply(z_d, scale, shift)
print("bn_relu_apply", rel(nchw(a_d), a_ref.detach()))
dz, sums = ops.bn_bwd(da.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev), a_d, z_d, mean.to(dev), rstd.to(dev), gamma.to(dev), 2 * 13 * 17)
print("bn_bwd dz", rel(nchw(dz), zt.grad), "dgamma", rel(sums[C:], bn.weight.grad), "dbeta", rel(sums[:C], bn.bias.grad))

# dgrad conv via packed dgrad weights
cin, cout = 128, 64
xw = bf(torch.randn(2, cin, 13, 17, generator=g)).requires_grad_(True)
wt = bf(torch.randn(cout, cin, 3, 3, generator=g) * 0.05)
zz = F.conv2d(xw, wt, padding=1)
dzz = bf(torch.randn_like(zz))
zz.backward(dzz)
wd = ops.pack_weight(1, wt.to(dev), cout, cin)
dx = ops.conv(dzz.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev), wd, cin, 9)
print("dgrad", rel(nchw(dx), xw.grad))

# maxpool bwd
ap = bf(torch.rand(2, 64, 12, 15, generator=g)).requires_grad_(True)
pp = F.max_pool2d(ap, 2)
dp = bf(torch.randn_like(pp)); dsk = bf(torch.randn(2, 64, 12, 15, generator=g))
pp.backward(dp)
df = ops.maxpool_bwd(ap.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev), dp.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev),
                     dsk.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev))
print("maxpool_bwd", rel(nchw(df), bf(ap.grad + dsk)))

# convT fwd / dgrad / wgrad
ci, co, hs, ws = 256, 128, 6, 7
xi = bf(torch.randn(2, ci, hs, ws, generator=g)).requires_grad_(True)
wtt = bf(torch.randn(ci, co, 2, 2, generator=g) * 0.05).requires_grad_(True)
bt = torch.randn(co, generator=g) * 0.1
uu = F.conv_transpose2d(xi, wtt, bt, stride=2)
Hf, Wf = 2 * hs + 1, 2 * ws + 1
du_full = bf(torch.randn(2, co, Hf, Wf, generator=g))
uu.backward(du_full[:, :, :2 * hs, :2 * ws])
dufd = du_full.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)
wdt = ops.pack_weight(3, wtt.detach().to(dev), co, ci)
din = ops.convt_dgrad(dufd, (0, 0), wdt, ci, hs, ws)
print("convT dgrad", rel(nchw(din), xi.grad))
gw = torch.empty(ci, co, 2, 2, device=dev)
ops.convt_wgrad(xi.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev), dufd, (0, 0), gw)
print("convT wgrad", rel(gw, wtt.grad))
wf = ops.pack_weight(2, wtt.detach().to(dev), co, ci)
uf = ops.conv(xi.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev), wf, co, ntaps=1, groups=4,
              scale=ops.ones(dev, 4 * co), shift=bt.repeat(4).to(dev))
print("convT fwd", rel(nchw(uf), uu.detach()))
# head bwd
al = bf(torch.rand(2, 64, 9, 11, generator=g)).requires_grad_(True)
wh = torch.randn(2, 64, 1, 1, generator=g).requires_grad_(True); bh = torch.randn(2, generator=g).requires_grad_(True)
yy = F.conv2d(al, wh, bh); dyy = torch.randn_like(yy); yy.backward(dyy)
dwh = torch.zeros(2, 64, device=dev); dbh = torch.zeros(2, device=dev)
dal = ops.head_bwd(al.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev), dyy.to(dev), wh.detach().reshape(2, 64).to(dev), dwh, dbh)
print("head_bwd da", rel(nchw(dal), al.grad), "dw", rel(dwh, wh.grad.reshape(2, 64)), "db", rel(dbh, bh.grad))
