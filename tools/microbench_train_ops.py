"""Achieved HBM GB/s of the memory-bound training operators at the layer sizes of the G2 training step
(batch 32): python tools/microbench_train_ops.py [batch]  ->  one line per (operator, layer size)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gelslim_depth_b200.train import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
BF = torch.bfloat16


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


class BN:
    def __init__(self, C):
        self.num_features = C
        self.weight = torch.rand(C, device=dev) + 0.5
        self.bias = torch.randn(C, device=dev)
        self.running_mean = torch.zeros(C, device=dev)
        self.running_var = torch.ones(C, device=dev)
        self.momentum = 0.1
        self.eps = 1e-5


rows = []
for (H, W, C) in [(320, 427, 64), (160, 213, 128), (80, 106, 256), (40, 53, 512), (20, 26, 1024)]:
    z = torch.randn(B, H, W, C, device=dev).to(BF)
    da = torch.randn(B, H, W, C, device=dev).to(BF)
    scale = torch.rand(C, device=dev) + 0.5
    shift = torch.randn(C, device=dev) * 0.1
    mean = torch.randn(C, device=dev) * 0.1
    rstd = torch.rand(C, device=dev) + 0.5
    gamma = torch.rand(C, device=dev) + 0.5
    nbytes = z.numel() * 2
    npix = B * H * W

    def rec(name, ms, passes):
        gbs = passes * nbytes / ms / 1e6
        rows.append({"op": name, "shape": [B, H, W, C], "ms": round(ms, 4), "GBps": round(gbs, 1)})
        print(f"{name:22s} {H}x{W}x{C:5d}  {ms:8.4f} ms  {gbs:8.1f} GB/s", flush=True)

    rec("bn_relu_apply", timeit(lambda: ops.bn_relu_apply(z, scale, shift, pool=False)), 2)
    rec("bn_relu_apply+pool", timeit(lambda: ops.bn_relu_apply(z, scale, shift, pool=True)), 2.25)
    rec("bn_bwd(reduce+apply)", timeit(lambda: ops.bn_bwd(da, scale, shift, z, mean, rstd, gamma, npix)), 5)
    sums = torch.zeros(2 * C, device=dev)
    import ctypes as Cc
    from gelslim_depth_b200._lib import lib, check
    st = ops._st(dev)
    rec("bn_bwd_reduce", timeit(lambda: check(lib.gsd_op_bn_bwd_reduce(ops._p(da), ops._p(scale), ops._p(shift), ops._p(z), ops._p(mean),
                                                                       ops._p(rstd), npix, C, ops._p(sums), st), "r")), 2)
    dz = torch.empty_like(z)
    rec("bn_bwd_apply", timeit(lambda: check(lib.gsd_op_bn_bwd_apply(ops._p(da), ops._p(scale), ops._p(shift), ops._p(z), ops._p(mean),
                                                                     ops._p(rstd), ops._p(gamma), ops._p(sums), float(npix), npix, C,
                                                                     ops._p(dz), st), "a")), 3)
    rec("channel_sum", timeit(lambda: ops.channel_sum(da)), 1)
    if H % 2 == 0 or True:
        a = torch.relu(torch.randn(B, H, W, C, device=dev)).to(BF)
        dpool = torch.randn(B, H // 2, W // 2, C, device=dev).to(BF)
        rec("maxpool_bwd(+skip)", timeit(lambda: ops.maxpool_bwd(a, dpool, da)), 3.25)
    if C == 64:
        w = torch.randn(2, 64, device=dev) * 0.1
        bias = torch.zeros(2, device=dev)
        dy = torch.randn(B, 2, H, W, device=dev)
        dw = torch.zeros(2, 64, device=dev)
        db = torch.zeros(2, device=dev)
        a = torch.relu(torch.randn(B, H, W, C, device=dev)).to(BF)
        rec("head_bwd", timeit(lambda: ops.head_bwd(a, dy, w, dw, db)), 2 + 2 * 4 / 128)
        rec("head_fwd", timeit(lambda: ops.head_fwd(a, w, bias)), 1 + 2 * 4 / 128)
    del z, da

n = 31_040_000
p, g, m, v, sh = (torch.randn(n, device=dev) for _ in range(5))
v.abs_()
ms = timeit(lambda: ops.adam_ema(p, g, m, v, sh, 1e-3, (0.9, 0.999), 1e-8, 1e-6, 5, 0.995, 5))
print(f"adam_ema 31.04M  {ms:.4f} ms  {9 * 4 * n / ms / 1e6:.1f} GB/s")
rows.append({"op": "adam_ema", "n": n, "ms": round(ms, 4), "GBps": round(9 * 4 * n / ms / 1e6, 1)})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/microbench_train_ops.json", "w"), indent=1)
