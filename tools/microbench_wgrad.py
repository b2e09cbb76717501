"""Per-layer timing of the conv weight-gradient kernel at the G2 training-step shapes (batch 32):
python tools/microbench_wgrad.py [batch]   ->  TFLOP/s per layer + total ms."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gelslim_depth_b200._lib import lib, check
from gelslim_depth_b200.train import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
BF = torch.bfloat16
# (name, H, W, C0, C1, Cout)
LAYERS = [("inc.0", 320, 427, 16, 0, 64), ("inc.3", 320, 427, 64, 0, 64), ("down.0.0", 160, 213, 64, 0, 128),
          ("down.0.3", 160, 213, 128, 0, 128), ("down.1.0", 80, 106, 128, 0, 256), ("down.1.3", 80, 106, 256, 0, 256),
          ("down.2.0", 40, 53, 256, 0, 512), ("down.2.3", 40, 53, 512, 0, 512), ("down.3.0", 20, 26, 512, 0, 1024),
          ("down.3.3", 20, 26, 1024, 0, 1024), ("up.0.conv.0", 40, 53, 512, 512, 512), ("up.0.conv.3", 40, 53, 512, 0, 512),
          ("up.1.conv.0", 80, 106, 256, 256, 256), ("up.1.conv.3", 80, 106, 256, 0, 256), ("up.2.conv.0", 160, 213, 128, 128, 128),
          ("up.2.conv.3", 160, 213, 128, 0, 128), ("up.3.conv.0", 320, 427, 64, 64, 64), ("up.3.conv.3", 320, 427, 64, 0, 64)]


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


tot_ms = tot_fl = 0.0
rows = []
for name, H, W, C0, C1, Co in LAYERS:
    x0 = torch.randn(B, H, W, C0, device=dev).to(BF)
    x1 = torch.randn(B, H, W, C1, device=dev).to(BF) if C1 else None
    dz = torch.randn(B, H, W, Co, device=dev).to(BF)
    dwk = torch.zeros(Co, 9, C0 + C1, device=dev)
    st = ops._st(dev)

    def run():
        check(lib.gsd_op_wgrad3x3_bf16(ops._p(x0), C0, ops._p(x1), C1, H if C1 else 0, W if C1 else 0, 0, 0, ops._p(dz), Co, B, H, W,
                                       ops._p(dwk), 0, st), "wgrad")
    ms = timeit(run)
    cin = 6 if C0 == 16 else C0 + C1
    fl = 2.0 * B * H * W * Co * 9 * cin
    tot_ms += ms
    tot_fl += fl
    rows.append({"layer": name, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)})
    print(f"{name:12s} {H}x{W} {C0 + C1:5d}->{Co:5d}  {ms:7.4f} ms  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
    del x0, x1, dz
print(f"total {tot_ms:.3f} ms  {tot_fl / tot_ms / 1e9:.1f} TFLOP/s")
rows.append({"layer": "total", "ms": round(tot_ms, 3), "tflops": round(tot_fl / tot_ms / 1e9, 1)})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/microbench_wgrad%s.json" % os.environ.get("TAG", ""), "w"), indent=1)
