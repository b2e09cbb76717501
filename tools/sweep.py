"""BASELINE configs[4]: batch-sharded inference sweep.  One process per GPU (plain `python tools/sweep.py` for one GPU,
`python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py` for N): every rank runs the
same plan on its own shard of the batch with no data-path collective; the time of a point is the max over ranks (NCCL is
used for the barrier and that max only).  Device-resident frames, plus (batch 64 per GPU) the blocking and the
rotating-slot host pipelines.  Rank 0 prints one JSON line per total batch."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gelslim_depth_b200.models.unet import UNet
from gelslim_depth_b200.engine import make_prepost


def cpu_path_fps(seconds=12.0):
    """the reference algorithm on this box's host cores (frames/s at batch 1 and batch 8): bench.py's `cpu_baseline` leg, the one
    place outside tests/ that may execute oracle/"""
    import time
    import bench
    torch.set_num_threads(os.cpu_count() or 1)
    out = {"cores": os.cpu_count()}
    for b in (1, 8):
        step = bench._cpu_step_fn(torch, b)
        ts, t_all = [], time.perf_counter()
        for i in range(6):
            t0 = time.perf_counter()
            step()
            if i:
                ts.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all > seconds / 2 and len(ts) >= 2:
                break
        out[f"batch{b}_frames_per_s"] = b / sorted(ts)[len(ts) // 2]
    return out


def run(per_gpu=(1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024), host_chunks=(16,), max_resident=256):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        per_gpu = tuple(b for b in per_gpu if b * world <= 1024)

    def maxms(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    torch.manual_seed(0)
    net = UNet(6, 2).to(dev).eval()
    H, W = 320, 427
    base = torch.randint(0, 256, (1, 6, H, W), dtype=torch.uint8).float().to(dev)
    pp = make_prepost(6, (H, W), (H, W), use_diff=True, in_scale=[1 / 255.0], out_scale=1.9180814027786255 / -0.9,
                      out_shift=-1.9180814027786255)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cpu = cpu_path_fps() if rank == 0 else None
    for B in per_gpu:
        # more than `max_resident` frames per GPU run as consecutive sub-batches through one plan (22 GB of activations
        # per 64 frames: 1024 resident frames would not fit 180 GB)
        Bp = min(B, max_resident)
        parts = B // Bp
        x = torch.randint(0, 256, (Bp, 6, H, W), dtype=torch.uint8,
                          generator=torch.Generator().manual_seed(100 + rank)).float().to(dev)
        y = torch.empty(Bp, 2, H, W, device=dev)
        plan = net.plan_for(Bp, H, W, dev)
        packed = net.packed_weights(plan)
        for _ in range(3):
            plan.forward(x, base, pp, y, packed)
        iters = max(3, min(100, int(1000 / B)))
        barrier()
        e0.record()
        for _ in range(iters * parts):
            plan.forward(x, base, pp, y, packed)
        e1.record()
        barrier()
        ms = maxms(e0.elapsed_time(e1)) / iters
        fps = world * B / ms * 1e3
        rec = {"n_gpus": world, "batch_total": world * B, "batch_per_gpu": B, "resident_sub_batch": Bp, "ms_per_batch": ms, "frames_per_s": fps,
               "tensor_frac_of_sustained": fps / world * 200.117 / 1e3 / 1386.1,
               "cpu_path": cpu, "speedup_vs_cpu_path": fps / max(cpu["batch1_frames_per_s"], cpu["batch8_frames_per_s"]) if cpu else None}
        if B == 64:
            xh, yh = x.cpu().pin_memory(), [torch.empty(B, 2, H, W).pin_memory() for _ in range(2)]
            xd, yd = [x, torch.empty_like(x)], [y, torch.empty_like(y)]
            for ch in host_chunks:
                plan.set_chunk(ch, first=ch // 2, last=ch // 2)
                for _ in range(2):
                    plan.forward_host(xh, base, pp, yh[0], x, y, packed)
                barrier()
                e0.record()
                for _ in range(5):
                    plan.forward_host(xh, base, pp, yh[0], x, y, packed)
                e1.record()
                barrier()
                rec[f"e2e_blocking_fps_chunk{ch}"] = world * B * 5 / (maxms(e0.elapsed_time(e1)) / 1e3)
            plan.set_chunk(B)

            def pipelined(n=10):
                for k in range(n):
                    plan.forward_host_async(xh, base, pp, yh[k % 2], xd[k % 2], yd[k % 2], packed, slot=k % 2)
                plan.host_wait(0)
                plan.host_wait(1)
            pipelined(4)
            barrier()
            e0.record()
            pipelined(10)
            e1.record()
            barrier()
            rec["e2e_rotating_slots_fps"] = world * B * 10 / (maxms(e0.elapsed_time(e1)) / 1e3)
            del xh, yh, xd, yd
        if rank == 0:
            print(json.dumps(rec), flush=True)
        del plan, x, y
        net._plans.clear()
        torch.cuda.empty_cache()
    if world > 1:
        import threading
        t = threading.Timer(20.0, lambda: os._exit(0))      # a blocking teardown must not hang the sweep
        t.daemon = True
        t.start()
        dist.destroy_process_group()
        t.cancel()


if __name__ == "__main__":
    run()
