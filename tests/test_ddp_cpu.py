"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: the bucket planner; the library's own backward driven as
a DRY RUN (gsd_debug_train_plan_create: the real step structure, nothing enqueued), whose gsd_bucket_cb callbacks feed the
BucketReducer FusedTrainer uses (here with gloo's all_reduce as the launcher); the 1/world gradient scale; the
construction-time broadcast of rank 0's model; and the frame sharding of batch-sharded inference."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _unet_entries():
    """(name, offset, numel) of UNet(6,2)'s 64 parameters in arena order, without instantiating 31 M floats"""
    dims = [64, 128, 256, 512, 1024]
    sizes = []

    def double(a, b):
        sizes.extend([b * a * 9, b, b, b * b * 9, b, b])

    double(6, dims[0])
    for lo, hi in zip(dims[:-1], dims[1:]):
        double(lo, hi)
    for hi, lo in zip(dims[:0:-1], dims[-2::-1]):
        sizes.extend([hi * (hi // 2) * 4, hi // 2])
        double(hi, lo)
    sizes.extend([2 * 64, 2])
    entries, off = [], 0
    for i, k in enumerate(sizes):
        entries.append((i, off, k))
        off += k
    return entries, off


def test_bucket_plan_of_the_full_unet():
    sys.path.insert(0, ROOT)
    from gelslim_depth_b200.train.engine import plan_buckets
    entries, total = _unet_entries()
    assert len(entries) == 64 and total > 31_000_000
    buckets, bucket_of = plan_buckets(entries, 25 << 20, first_bucket_bytes=4 << 20, tail_bucket_bytes=2 << 20)
    # contiguous cover of the arena, formed from its END (backward order); every parameter in exactly one bucket
    assert sorted(k for b in buckets for k in b["params"]) == list(range(64))
    assert buckets[0]["hi"] == total and buckets[-1]["lo"] == 0
    for a, b in zip(buckets[:-1], buckets[1:]):
        assert b["hi"] == a["lo"]
    mb = [(b["hi"] - b["lo"]) * 4 / 2 ** 20 for b in buckets]
    assert mb[0] <= 8, mb          # the first all-reduce starts after the head + last decoder blocks
    assert mb[-1] <= 2.5, mb       # the all-reduce that nothing can hide (inc.*, down.0.*) is short
    assert 63 in buckets[0]["params"] and 0 in buckets[-1]["params"]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    from gelslim_depth_b200 import _lib
    from gelslim_depth_b200._lib import lib
    from gelslim_depth_b200.train.engine import BucketReducer, broadcast_module_state, plan_buckets
    # a small U-Net geometry: the library lays out its 28 parameters and walks its real backward structure (dry run)
    g = _lib.Geometry()
    g.batch, g.in_channels, g.height, g.width, g.n_classes, g.n_dims = 2, 3, 24, 29, 1, 3
    for i, d in enumerate((64, 128, 256)):
        g.dims[i] = d
    g.dtype, g.mode = _lib.DTYPE_BF16, _lib.MODE_TRAIN
    h = C.c_void_p()
    assert lib.gsd_debug_train_plan_create(C.byref(h), C.byref(g)) == 0, lib.gsd_last_error()
    n = lib.gsd_train_plan_num_params(h)
    numel = (C.c_longlong * n)()
    assert lib.gsd_train_plan_param_numel(h, numel, n) == n == 6 * 3 + 8 * 2 + 2
    entries, off = [], 0
    for i, k in enumerate(numel):
        entries.append((i, off, int(k)))
        off += int(k)
    buckets, bucket_of = plan_buckets(entries, bucket_bytes=1 << 20, first_bucket_bytes=1 << 16, tail_bucket_bytes=1 << 16)
    assert len(buckets) >= 3 and sorted(k for b in buckets for k in b["params"]) == list(range(n))
    assert buckets[0]["hi"] == off and buckets[-1]["lo"] == 0
    for a, b in zip(buckets[:-1], buckets[1:]):
        assert b["hi"] == a["lo"]
    nb = len(buckets)
    assert lib.gsd_train_plan_set_buckets(h, nb, (C.c_int * n)(*[bucket_of[i] for i in range(n)]),
                                          (C.c_longlong * nb)(*[b["lo"] for b in buckets]),
                                          (C.c_longlong * nb)(*[b["hi"] for b in buckets])) == 0
    gen = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(off, generator=gen)            # this rank's "gradient arena"
    mine = flat.clone()
    launches = []

    def launch(lo, hi):
        launches.append((lo, hi))
        dist.all_reduce(flat[lo:hi])

    red = BucketReducer(nb, launch)
    red.begin_step()
    cb = _lib.BUCKET_CB(lambda user, bucket, lo, hi, main, side: red.on_bucket(bucket, lo, hi))
    assert lib.gsd_backward(h, C.c_void_p(16), None, cb, None) == 0, lib.gsd_last_error()      # dry run: only the callbacks happen
    ok = red.all_fired() and red.fired == list(range(nb))            # buckets complete in arena-END-first order
    ok = ok and launches[0][1] == off and launches[-1][0] == 0
    flat *= 1.0 / world                                   # grad_scale of gsd_adam_ema_step
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    want = sum(gathered) / world
    ok = ok and torch.allclose(flat, want, rtol=1e-6, atol=1e-6)
    try:
        red.on_bucket(0, 0, 1)
        ok = False                                        # a second report of one bucket in a step must raise
    except RuntimeError:
        pass
    lib.gsd_train_plan_destroy(h)
    # construction-time broadcast (DistributedDataParallel semantics): replicas built from different seeds end up
    # with rank 0's parameters AND BatchNorm buffers (float and int64)
    torch.manual_seed(1234 + rank)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8))
    with torch.no_grad():
        net[1].running_mean.normal_()
        net[1].num_batches_tracked.fill_(7 + rank)
    flat_p = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    broadcast_module_state(net, flat_p)
    state = torch.cat([flat_p, net[1].running_mean, net[1].running_var, net[1].num_batches_tracked.float().reshape(1)])
    lo_, hi_ = state.clone(), state.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    ok = ok and torch.equal(lo_, hi_) and int(net[1].num_batches_tracked) == 7
    # batch-sharded inference: frames [rank*B, (rank+1)*B) per rank, throughput adds up (bench.py weak scaling)
    frames = torch.tensor([64.0 * 5])
    dist.all_reduce(frames)
    q.put((rank, ok, float(frames)))
    dist.destroy_process_group()


def test_bucketed_allreduce_broadcast_and_sharding_gloo():
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
