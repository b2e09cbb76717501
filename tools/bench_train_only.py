"""Training step only (BASELINE configs[3]) with more timed steps than bench.py's default: python tools/bench_train_only.py [steps] [batch]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
r = bench.bench_train(torch, None, dev, 0, 1, batch, steps)
r.pop("losses")
print(json.dumps(r))
