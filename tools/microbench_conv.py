"""Time single tcgen05 conv launches (tuning aid).  python tools/microbench_conv.py cin cout H W B [bn] [pool]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gelslim_depth_b200.engine import conv_op

TAPS3 = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]


def run(cin, cout, H, W, B, bn=0, pool=False, iters=10, taps=TAPS3):
    d = torch.device("cuda:0")
    x = torch.randn(B, H, W, cin, device=d).to(torch.bfloat16)
    w = (torch.randn(cout, len(taps) * cin, device=d) * 0.05).to(torch.bfloat16)
    sc, sh = torch.ones(cout, device=d), torch.zeros(cout, device=d)
    for _ in range(3):
        conv_op(x, w, sc, sh, taps, relu=True, block_n=bn, pool=pool)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        conv_op(x, w, sc, sh, taps, relu=True, block_n=bn, pool=pool)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * B * H * W * cout * len(taps) * cin
    print(f"cin={cin} cout={cout} {H}x{W} B={B} bn={bn} pool={pool} tile={os.environ.get('GSD_FORCE_TILE','auto')} "
          f"taps={len(taps)}: {ms:.4f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:6]]
    bn = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    pool = len(sys.argv) > 7 and sys.argv[7] == "pool"
    ntaps = int(sys.argv[8]) if len(sys.argv) > 8 else 9
    run(*a, bn=bn, pool=pool, taps=TAPS3[:ntaps] if ntaps < 9 else TAPS3)
