"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: gradient bucketing / all-reduce / 1/world scaling
used by FusedTrainer, and the frame sharding of batch-sharded inference."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gelslim_depth_b200.train.engine import plan_buckets
    sizes = [1728, 64, 64, 36864, 64, 64, 73728, 128, 128, 147456, 2097152, 512, 128, 2]     # mini parameter list
    entries, off = [], 0
    for i, k in enumerate(sizes):
        entries.append((i, off, k))
        off += k
    buckets, bucket_of = plan_buckets(entries, bucket_bytes=1 << 20)
    # every parameter in exactly one bucket, buckets contiguous, formed from the END of the arena (backward order)
    assert sorted(k for b in buckets for k in b["params"]) == list(range(len(sizes)))
    assert buckets[0]["hi"] == off and buckets[-1]["lo"] == 0
    for a, b in zip(buckets[:-1], buckets[1:]):
        assert b["hi"] == a["lo"]
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(off, generator=g)
    mine = flat.clone()
    pending = [len(b["params"]) for b in buckets]
    for i in reversed(range(len(sizes))):               # gradients become ready in reverse parameter order
        bi = bucket_of[i]
        pending[bi] -= 1
        if pending[bi] == 0:
            dist.all_reduce(flat[buckets[bi]["lo"]:buckets[bi]["hi"]])
    flat *= 1.0 / world                                   # grad_scale of gsd_op_adam_ema
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    want = sum(gathered) / world
    ok = torch.allclose(flat, want, rtol=1e-6, atol=1e-6)
    # batch-sharded inference: frames [rank*B, (rank+1)*B) per rank, throughput adds up (bench.py weak scaling)
    frames = torch.tensor([64.0 * 5])
    dist.all_reduce(frames)
    q.put((rank, ok, float(frames)))
    dist.destroy_process_group()


def test_bucketed_allreduce_and_sharding_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert all(abs(f - 640.0) < 1e-6 for _, _, f in res)
