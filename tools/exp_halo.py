"""Halo-kernel experiment: numerics for both descriptor base-offset modes, then timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from gelslim_depth_b200.engine import conv3x3_halo_op, conv_op

d = torch.device("cuda:0")


def pack_w3(w):
    O, I = w.shape[:2]
    return w.reshape(O, I, 9).permute(0, 2, 1).reshape(O, 9 * I).contiguous().to(torch.bfloat16)


def numerics(cin, cout, h, w, b, mode, pool=False, cin1=0):
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(b, cin + cin1, h, w, generator=g).to(torch.bfloat16).float()
    wt = (torch.randn(cout, cin + cin1, 3, 3, generator=g) * (2.0 / (9 * (cin + cin1))) ** 0.5).to(torch.bfloat16).float()
    sc, sh = 0.5 + torch.rand(cout, generator=g), 0.3 * torch.randn(cout, generator=g)
    ref = torch.relu(F.conv2d(x, wt, padding=1) * sc[None, :, None, None] + sh[None, :, None, None])
    xs = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(d)
    s0 = xs[..., :cin].contiguous()
    s1 = xs[..., cin:].contiguous() if cin1 else None
    # weight K order: tap-major then [src0 channels, src1 channels]
    r = conv3x3_halo_op(s0, pack_w3(wt).to(d), sc.to(d), sh.to(d), relu=True, src1=s1, pool=pool, base_off_mode=mode)
    torch.cuda.synchronize()
    out = (r[0] if pool else r).permute(0, 3, 1, 2).float().cpu()
    err = (out - ref).abs()
    bad = int((err > ref.abs() * 2 ** -7 + 1e-3).sum())
    msg = f"mode={mode} cin={cin}+{cin1} cout={cout} {h}x{w} b={b} pool={pool}: bad={bad}/{err.numel()} maxerr={float(err.max()):.4g}"
    if pool:
        want = F.max_pool2d(out, 2)
        msg += f" pool_exact={bool(torch.equal(r[1].permute(0, 3, 1, 2).float().cpu(), want))}"
    print(msg, flush=True)
    return bad == 0


def timing(cin, cout, H, W, B, pool=False, iters=10):
    x = torch.randn(B, H, W, cin, device=d).to(torch.bfloat16)
    w = (torch.randn(cout, 9 * cin, device=d) * 0.05).to(torch.bfloat16)
    sc, sh = torch.ones(cout, device=d), torch.zeros(cout, device=d)
    for fn, name in ((lambda: conv3x3_halo_op(x, w, sc, sh, pool=pool), "halo"),
                     (lambda: conv_op(x, w, sc, sh, [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1)], pool=pool), "tap9")):
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
        print(f"{name}: cin={cin} cout={cout} {H}x{W} B={B} pool={pool}: {ms:.4f} ms {2.0*B*H*W*cout*9*cin/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    ok1 = numerics(64, 64, 19, 23, 2, 1)
    ok0 = numerics(64, 64, 19, 23, 2, 0)
    print("base_off_mode=1 correct:", ok1, "| base_off_mode=0 correct:", ok0, flush=True)
    if ok0:
        numerics(64, 64, 40, 53, 2, 0, pool=True)
        numerics(64, 128, 33, 20, 1, 0)
        numerics(128, 128, 21, 27, 2, 0, pool=True)
        numerics(256, 256, 16, 24, 3, 0)
        numerics(64, 64, 20, 26, 2, 0, cin1=64)
        numerics(128, 128, 10, 13, 1, 0, cin1=128)
        timing(64, 64, 320, 427, 16)
        timing(64, 64, 320, 427, 16, pool=True)
        timing(128, 64, 320, 427, 16)
        timing(64, 128, 160, 213, 16)
        timing(128, 128, 160, 213, 16)
        timing(256, 128, 160, 213, 16)
        timing(256, 256, 80, 106, 16)
        timing(512, 512, 40, 53, 16)
