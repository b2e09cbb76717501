"""Repeat the 200-step loss-curve run (tests/test_gpu_train.py::test_loss_curve_200_steps_vs_reference) N times and print the
deviation statistics of every run: a race would show up as an outlier.  python tests/stress_train_curve.py [runs]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
from gelslim_depth_b200.models.unet import UNet
from gelslim_depth_b200.train.engine import FusedTrainer
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
g = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "train_curve.pt"), weights_only=False)
ref = g["losses"]
smooth = lambda v: [sum(v[max(0, i - 9): i + 1]) / len(v[max(0, i - 9): i + 1]) for i in range(len(v))]
b = smooth(ref)
worst = 0.0
stats = []
for r in range(runs):
    torch.manual_seed(g["module_seed"])
    net = UNet(3, 1)
    net.load_state_dict(oracle.trainer_init_state_dict(net.state_dict(), seed=g["init_seed"]))
    net = net.to(dev).train()
    ft = FusedTrainer(net)
    X, T = g["X"].to(dev), g["T"].to(dev)
    losses = []
    for step in range(g["steps"]):
        idx = torch.arange(4) + 4 * (step % 4)
        losses.append(ft.step(X[idx], T[idx]))
    losses = [float(v) for v in torch.cat(losses).cpu()]
    bad = [i for i, v in enumerate(losses) if not (v == v and v > 0)]
    if bad:
        print(f"run {r}: non-finite / non-positive loss at steps {bad[:5]}", flush=True)
        continue
    a = smooth(losses)
    dist = [abs(math.log10(x) - math.log10(y)) for x, y in zip(a, b)]
    worst = max(worst, max(dist))
    med = lambda v: sorted(v[-20:])[10]
    p90 = sorted(dist)[int(0.9 * len(dist))]
    stats.append((max(dist), sum(dist) / len(dist), p90, med(losses) / med(ref), losses[-1] / ref[-1]))
    print(f"run {r}: first {losses[0]:.4e} final {losses[-1]:.3e} (ref {ref[-1]:.3e}) max dist {max(dist):.3f} at step {dist.index(max(dist))} mean {sum(dist) / len(dist):.3f} p90 {p90:.3f} median-last-20 ratio {med(losses) / med(ref):.3f}", flush=True)
print("worst max dist", worst)
for name, i in (("max dist", 0), ("mean dist", 1), ("p90 dist", 2), ("median-last-20 / ref", 3), ("last / ref", 4)):
    v = sorted(t[i] for t in stats)
    print(f"{name}: min {v[0]:.3f} median {v[len(v) // 2]:.3f} max {v[-1]:.3f}")
