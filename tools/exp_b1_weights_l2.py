"""Batch-1 deep layers: one conv launch with its weights L2-resident (back-to-back launches) vs evicted (256 MB written
between launches, what a whole forward does to the 126 MB L2).  Decides whether keeping weights L2-resident pays."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gelslim_depth_b200.engine import conv_op

TAPS3 = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
d = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for cin, cout, H, W in ((512, 1024, 20, 26), (1024, 1024, 20, 26), (1024, 512, 40, 53), (512, 512, 40, 53), (256, 512, 40, 53),
                        (512, 256, 80, 106), (256, 256, 80, 106), (128, 128, 160, 213), (64, 64, 320, 427)):
    x = torch.randn(B, H, W, cin, device=d).to(torch.bfloat16)
    w = (torch.randn(cout, 9 * cin, device=d) * 0.05).to(torch.bfloat16)
    sc, sh = torch.ones(cout, device=d), torch.zeros(cout, device=d)
    def graph_ms(body, n=10):
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            body(); body()
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n):
                body()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / n)
        return statistics.median(ts)
    conv = lambda: conv_op(x, w, sc, sh, TAPS3, relu=True)
    t_warm = graph_ms(conv)
    t_flush = graph_ms(lambda: flush.fill_(1))
    t_cold = graph_ms(lambda: (flush.fill_(1), conv())) - t_flush
    t_xwarm = graph_ms(lambda: (flush.fill_(1), x.add_(0), conv())) - graph_ms(lambda: (flush.fill_(1), x.add_(0)))
    print(f"{cin:5d}->{cout:5d} {H}x{W} B={B}: weights+input in L2 {t_warm:6.1f} us   both evicted {t_cold:6.1f} us   only weights evicted {t_xwarm:6.1f} us", flush=True)
