"""Runs the ctypes stub printed in INTEGRATION.md (section B) as is and checks it against the oracle."""
import os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
src = open(os.path.join(ROOT, "INTEGRATION.md")).read()
block = re.search(r"```python\nimport ctypes as C, torch\n(.*?)```", src, re.S).group(0)
code = block[len("```python\n"):-3].replace('C.CDLL("libgsd_b200.so")', 'C.CDLL("%s")' % os.path.join(ROOT, "gelslim_depth_b200", "libgsd_b200.so"))
ns = {}
exec(code, ns)
import torch, oracle
from gelslim_depth_b200.models.unet import UNet
torch.manual_seed(0)
net = UNet(6, 2)
sd = oracle.conditioned_state_dict(net.state_dict(), seed=5)
net.load_state_dict(sd)
net = net.cuda().eval()
x = torch.rand(2, 6, 48, 59).cuda()
y = ns["unet_forward_b200"](net, x)
torch.cuda.synchronize()
ref = oracle.unet_forward(sd, x.cpu())
print("stub rel_l2", float((y.cpu() - ref).norm() / ref.norm()))
