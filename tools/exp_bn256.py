"""Experiment: halo conv with N = 256 UMMAs (block_n=256) vs the default N = 128 x MT = 2; numerics + timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from gelslim_depth_b200.engine import conv3x3_halo_op
d = torch.device("cuda:0")

def pack_w3(w):
    O, I = w.shape[:2]
    return w.reshape(O, I, 9).permute(0, 2, 1).reshape(O, 9 * I).contiguous().to(torch.bfloat16)

def numerics(cin, cout, h, w, b, bn):
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(b, cin, h, w, generator=g).to(torch.bfloat16).float()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5).to(torch.bfloat16).float()
    sc, sh = 0.5 + torch.rand(cout, generator=g), 0.3 * torch.randn(cout, generator=g)
    ref = torch.relu(F.conv2d(x, wt, padding=1) * sc[None, :, None, None] + sh[None, :, None, None])
    xs = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(d)
    r = conv3x3_halo_op(xs, pack_w3(wt).to(d), sc.to(d), sh.to(d), relu=True, block_n=bn)
    torch.cuda.synchronize()
    out = r.permute(0, 3, 1, 2).float().cpu()
    err = (out - ref).abs()
    bad = int((err > ref.abs() * 2 ** -7 + 1e-3).sum())
    print(f"numerics bn={bn} {cin}->{cout} {h}x{w} b={b}: bad={bad}/{err.numel()} maxerr={float(err.max()):.4g}", flush=True)

def timing(cin, cout, H, W, B, bn, iters=10):
    x = torch.randn(B, H, W, cin, device=d).to(torch.bfloat16)
    w = (torch.randn(cout, 9 * cin, device=d) * 0.05).to(torch.bfloat16)
    sc, sh = torch.ones(cout, device=d), torch.zeros(cout, device=d)
    fn = lambda: conv3x3_halo_op(x, w, sc, sh, block_n=bn)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"timing bn={bn} {cin}->{cout} {H}x{W} B={B}: {ms:.4f} ms {2.0 * B * H * W * cout * 9 * cin / ms / 1e9:.1f} TFLOP/s", flush=True)

numerics(64, 256, 40, 53, 8, 256)
numerics(128, 512, 33, 20, 20, 256)
for (ci, co, H, W) in [(256, 256, 80, 106), (128, 256, 80, 106), (512, 512, 40, 53), (256, 512, 40, 53), (512, 256, 80, 106), (1024, 512, 40, 53)]:
    for bn in (0, 256):
        timing(ci, co, H, W, 64, bn)
