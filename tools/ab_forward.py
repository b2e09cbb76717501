"""A/B of plan variants inside ONE process on ONE box: per-launch medians of gsd_forward_profiled, variants interleaved
round-robin so clock / thermal drift cancels.  Variants are environment settings that the library reads when a plan
binds (GSD_NO_BIAS_MMA, GSD_NO_FUSED_PROLOGUE, GSD_FIRST_DBG, ...).

    python tools/ab_forward.py 64 "A:" "B:GSD_NO_BIAS_MMA=1" "C:GSD_NO_FUSED_PROLOGUE=1,GSD_NO_BIAS_MMA=1"
"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gelslim_depth_b200.engine import Plan, make_prepost  # noqa: E402
from gelslim_depth_b200.models.unet import UNet  # noqa: E402

H, W = 320, 427
B = int(sys.argv[1])
variants = []
for spec in sys.argv[2:]:
    name, _, envs = spec.partition(":")
    variants.append((name, dict(kv.split("=") for kv in envs.split(",") if kv)))
all_keys = sorted({k for _, e in variants for k in e})
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = UNet(6, 2).to(dev).eval()
x = torch.randint(0, 256, (B, 6, H, W), device=dev).float()
base = torch.randint(0, 256, (1, 6, H, W), device=dev).float()
y = torch.empty(B, 2, H, W, device=dev)
pp = make_prepost(6, (H, W), (H, W), use_diff=True, in_scale=[1 / 255.0], out_scale=-2.13, out_shift=-1.9)
plans = []
for name, env in variants:
    for k in all_keys:
        os.environ.pop(k, None)
    os.environ.update(env)
    plan = Plan(B, 6, H, W, 2, net.layer_dimensions, dev)
    packed = torch.empty(plan.packed_bytes, dtype=torch.uint8, device=dev)
    plan.pack([p.detach() for p in net.parameters()], net._bn_buffers(), packed)
    plan.forward(x, base, pp, y, packed)           # binds under this environment
    torch.cuda.synchronize()
    if plans and not torch.equal(y, y_first):
        print(f"!! variant {name}: output differs from the first variant (max abs {(y - y_first).abs().max().item():.3e})")
    if not plans:
        y_first = y.clone()
    plans.append((name, env, plan, packed))
rows = {name: [] for name, *_ in plans}
steps = {name: [] for name, *_ in plans}
for rep in range(12):
    for name, env, plan, packed in plans:
        for k in all_keys:
            os.environ.pop(k, None)
        os.environ.update(env)                      # run-time switches (GSD_NO_FUSED_PROLOGUE) are read per call
        rows[name].append(plan.forward_profiled(x, base, pp, y, packed))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            plan.forward(x, base, pp, y, packed)
        e1.record()
        torch.cuda.synchronize()
        steps[name].append(e0.elapsed_time(e1) / 3)
for name, env, plan, _ in plans:
    r = rows[name]
    n = len(r[0])
    med = [statistics.median(rr[i][0] for rr in r) for i in range(n)]
    fused = n == 23
    names = (["first(fused)"] if fused else ["prologue", "inc.0"]) + ["inc.3"] + [f"down.{i}.{j}" for i in range(4) for j in (0, 3)] + \
            [f"up.{i}.{k}" for i in range(4) for k in ("up", "conv.0", "conv.3")] + ["tail"]
    print(f"== {name} {env}: step {statistics.median(steps[name]):.3f} ms, sum of launches {sum(med):.3f} ms")
    print("   " + "  ".join(f"{a}={m:.3f}" for a, m in zip(names, med)))
