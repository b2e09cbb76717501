"""ConvTranspose2d weight-gradient kernel at the four decoder levels (batch 32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gelslim_depth_b200.train import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
tot = 0.0
for (h, w, cin) in [(20, 26, 1024), (40, 53, 512), (80, 106, 256), (160, 213, 128)]:
    cout = cin // 2
    x = torch.randn(B, h, w, cin, device=dev).to(torch.bfloat16)
    du = torch.randn(B, 2 * h + (1 if h != 20 else 0), 2 * w + 1, cout, device=dev).to(torch.bfloat16)
    g = torch.zeros(cin, cout, 2, 2, device=dev)
    fn = lambda: ops.convt_wgrad(x, du, (0, 0), g)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tot += ms
    print(f"{cin}->{cout} {h}x{w}: {ms:.4f} ms (incl. grad zero fill) {2.0 * B * h * w * cin * cout * 4 / ms / 1e9:.0f} TFLOP/s")
print("total", round(tot, 4))
