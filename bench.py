#!/usr/bin/env python
"""bench.py -- U-Net frames/s @6x320x427 (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference algorithm on the host CPU cores

A "step" is one pass of the hot path (gsd_forward: input prologue -> 22 tcgen05 conv GEMMs -> 1x1 head)
over one batch of synthetic frame pairs.  Workload at every N: BASELINE.json configs[1] -- UNet(6,2),
bf16, batch 64 frames of 6x320x427 PER GPU (weak scaling; inference shards by frame with no collective).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, CIN, NCLS = 320, 427, 6, 2
GFLOP_PER_FRAME = 200.117          # 2*MAC over the 23 conv / transposed-conv layers (G2)
DIMS = [64, 128, 256, 512, 1024]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1386.1), d.get("bf16_tflops", 1634.5), d.get("hbm_gbs", 6541.5), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


def traffic_from_profile():
    """dram__bytes_read.sum + dram__bytes_write.sum of the 22 conv launches of one batch-64 forward, from the committed
    `ncu --set full` capture (profiles/r2_traffic.json); bytes per step, like `achieved` is FLOPs per step."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return {"dram_bytes_per_step": d["traffic_bytes_per_step"], "algorithmic_bytes_per_step": d["algorithmic_activation_bytes_per_step"] + d["weights_bytes"],
            "source": d["source"]}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: an NVML polling thread (every 5 ms; the timed region
    of the default run is ~120 ms, shorter than nvidia-smi's start-up), nvidia-smi -lms as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"),
               (0x80, "hw_power_brake_slowdown"))

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.handle, self.samples, self._stop = None, None, [], threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            phys = int(ids[index]) if ids and all(v.isdigit() for v in ids) and index < len(ids) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, mask))
            except Exception:
                pass
            time.sleep(0.005)

    def mark(self):
        """drop what was sampled so far: called at the start of the timed region (the sampler is started earlier so
        that NVML's first, slow calls are over by then)"""
        self.samples, self.lines = [], []

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1)
            n = self.nvml
            try:
                mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            except Exception:
                mx = None
            sm = [float(s) for s, _ in self.samples]
            mask = 0
            for _, m in self.samples:
                mask |= m
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": mx,
                    "reasons": sorted(name for bit, name in self.REASONS if mask & bit), "samples": len(sm), "source": "nvml 5 ms"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 50"}


def synthetic_frames(torch, batch, seed):
    """raw = randint(0,256) frames, base = randint(0,256) undeformed image."""
    g = torch.Generator().manual_seed(seed)
    raw = torch.randint(0, 256, (batch, CIN, H, W), generator=g, dtype=torch.uint8).float()
    base = torch.randint(0, 256, (1, CIN, H, W), generator=torch.Generator().manual_seed(1), dtype=torch.uint8).float()
    return raw, base


def _cpu_step_fn(torch, batch, seed=0):
    """one pass of the reference algorithm (oracle port of get_difference_image + normalisation + UNet.forward +
    depth de-normalisation) over `batch` synthetic 6x320x427 frame pairs on the host cores; weights are the reference
    constructor's random init (seed 0), drawn by the oracle without touching the product package"""
    import oracle
    sd = oracle.random_init_state_dict(CIN, NCLS, DIMS, seed=0)
    raw, base = synthetic_frames(torch, batch, seed)

    def step():
        with torch.no_grad():
            x = oracle.normalize_tactile_image(oracle.get_difference_image(raw, base), "0_255_to_0_1", 0.9, None)
            y = oracle.unet_forward(sd, x)
            return oracle.denormalize_depth_image(y, "min_max_to_0_-1", 0.9, (-1.9180814027786255, 0.0))
    return step


def cpu_reference_fps(torch, seconds_budget=24.0, threads=None):
    """The reference algorithm on the host cores, fp32, on a bounded sample of the batch-64 workload: single frame pairs
    and batches of 8 (the CPU path is faster per frame on a batch; the better of the two is the baseline)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    best, notes = None, []
    for batch, budget in ((1, 0.4 * seconds_budget), (8, 0.6 * seconds_budget)):
        step = _cpu_step_fn(torch, batch)
        times, t_all = [], time.perf_counter()
        for i in range(9):
            t0 = time.perf_counter()
            step()
            dt = time.perf_counter() - t0
            if i > 0:                       # first call = warm-up (oneDNN primitive creation)
                times.append(dt)
            if time.perf_counter() - t_all > budget and len(times) >= 2:
                break
        med = statistics.median(times)
        fps = batch / med
        notes.append(f"batch {batch}: {len(times)} timed fp32 forwards, median {med:.3f} s = {fps:.2f} frames/s")
        if best is None or fps > best:
            best = fps
    return best, threads, "oracle port on 6x320x427 frame pairs, 1 warm-up each; " + "; ".join(notes) + "; value = the better"


def bench_train(torch, dist, dev, rank, world, batch, steps):
    """train_unet.py:346-377 loop body on synthetic data: UNet(6,2) @ 6x320x427, bf16, `batch` samples per GPU,
    trainer init N(0, 0.01), Adam(1e-3, wd 1e-6), EMA 0.995; data-parallel gradient all-reduce over NCCL.
    world > 1: every rank builds its net from a DIFFERENT seed (the trainer broadcasts rank 0's model at construction,
    like DistributedDataParallel) and feeds different data; after the timed steps the replicas must hold bit-identical
    parameters / EMA shadow / loss-independent state (`ddp_in_sync`), and one extra backward checks that the all-reduced
    gradient arena equals the mean of the per-rank gradients."""
    from gelslim_depth_b200.models.unet import UNet
    from gelslim_depth_b200.train.engine import FusedTrainer
    torch.manual_seed(1000 * rank)
    net = UNet(CIN, NCLS, layer_dimensions=DIMS)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if "weight" in name:
                torch.nn.init.normal_(p, mean=0, std=0.01)          # train_unet.py:248-250
    net = net.to(dev).train()
    kw = {}
    if os.environ.get("GSD_DDP_BUCKET_MB"):            # tuning experiments only: "bucket,first,tail" in MB
        b, f, t = (float(v) for v in os.environ["GSD_DDP_BUCKET_MB"].split(","))
        kw = dict(bucket_bytes=int(b * 2 ** 20), first_bucket_bytes=int(f * 2 ** 20) or None, tail_bucket_bytes=int(t * 2 ** 20) or None)
    ft = FusedTrainer(net, use_graph=True, **kw)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand(batch, CIN, H, W, generator=g).to(dev)
    t = (-0.9 * torch.rand(batch, NCLS, H, W, generator=g)).to(dev)
    stream = torch.cuda.current_stream(dev)
    losses = []
    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(3):
        losses.append(ft.step(x, t))
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    if sampler:
        sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        losses.append(ft.step(x, t))
    e1.record(stream)
    torch.cuda.synchronize(dev)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    in_sync, grad_err = None, None
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt)
        # (1) replicas in sync: parameter / EMA checksums are bit-identical on every rank
        cs = ft.param_checksum()
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool(torch.equal(lo, hi)) and bool(torch.isfinite(cs).all())
        # (2) the bucketed, overlapped all-reduce delivers sum over ranks of the per-rank gradients
        g_red = ft.backward_only(x, t, reduce=True) / world
        g_loc = ft.backward_only(x, t, reduce=False)
        parts = [torch.empty_like(g_loc) for _ in range(world)]
        dist.all_gather(parts, g_loc)
        mean = torch.stack(parts).mean(0)
        grad_err = float((g_red - mean).abs().max() / (mean.abs().max() + 1e-30))   # backward uses fp32 atomics: not bit-reproducible
    buckets_mb = [round((b["hi"] - b["lo"]) * 4 / 2 ** 20, 1) for b in ft.buckets]
    train_launches = ft.plan.launches
    ft.close()                  # the captured graph holds NCCL nodes: free it while the process group is alive
    del ft
    sps = world * batch * steps / (ms / 1e3)
    sustained, burst = measured_peaks()[:2]
    tf = sps / world * 599.41 / 1e3
    return {"metric": "train_samples_per_s_6x320x427", "value": sps, "unit": "samples/s", "ms_per_step": ms / steps,
            "batch_per_gpu": batch, "steps": steps, "gflop_per_sample": 599.41, "tflops_per_gpu": tf,
            "tensor_frac_sustained": tf / sustained, "tensor_frac_burst": tf / burst,
            "ddp_in_sync": in_sync, "ddp_reduced_grad_rel_err": grad_err,
            "ddp_buckets_mb": buckets_mb, "launches_per_step": train_launches,
            "losses": [round(float(v), 6) for v in torch.cat(losses).cpu()], "clocks": clocks,
            "what": "fwd (train-mode BN) + MSE + bwd (dgrad/wgrad on tcgen05) + bucketed NCCL all-reduce + fused Adam/EMA, "
                    "whole step replayed as one CUDA graph"}


def workload_config(batch):
    """the same `config` on both arms (b200 / reference): BASELINE.json configs[1]"""
    return {"workload": "configs[1]: UNet(6,2) inference, batch %d frame pairs of 6x320x427 per GPU, difference image + "
                        "normalisation + depth de-normalisation inside the timed call" % batch,
            "geometry": "G2", "batch_per_gpu": batch,
            "l2": "inputs_larger_than_l2 (210 MB frames, 22 GB activations per step at batch 64)", "weights": "random init seed 0"}


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port -- the reference itself cannot
    travel to the GPU box) on all host threads; each step = `--cpu-batch` frame pairs (a bounded sample of the batch-64
    workload).  Imports nothing from the product package: no native library is loaded in this process."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cpu_batch, probe = args.cpu_batch, None
    if cpu_batch <= 0:
        # the CPU path's best operating point on this box: single frames or batches of 8 (oneDNN is usually FASTER per frame
        # on single frames here: 5.5 vs 4.4 frames/s on 16 threads), probed with one warm + one timed pass each
        probe = {}
        for b in (1, 8):
            fn = _cpu_step_fn(torch, b)
            fn()
            t = time.perf_counter()
            fn()
            probe[b] = b / (time.perf_counter() - t)
        cpu_batch = max(probe, key=probe.get)
    step = _cpu_step_fn(torch, cpu_batch)
    for _ in range(args.warmup):
        step()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t1
    fps = args.steps * cpu_batch / dt
    sample = (f"each step = {cpu_batch} 6x320x427 frame pair(s) (bounded sample of the batch-64 workload), fp32, "
              f"{threads} host threads" + (f"; batch chosen by a probe: {({k: round(v, 2) for k, v in probe.items()})} frames/s" if probe else ""))
    line = {"impl": "reference", "metric": "unet_frames_per_s_6x320x427", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.batch),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    emit(line)


_REAL_STDOUT = None


def shutdown(dist, world):
    """leave the process group; a teardown that blocks (peer already gone, captured collectives ...) must never hang the
    bench: the line is out, a watchdog ends the process"""
    if world <= 1:
        return
    import gc
    gc.collect()
    t = threading.Timer(20.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    try:
        dist.destroy_process_group()
    finally:
        t.cancel()


def emit(line: dict):
    """the ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was diverted to stderr"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def bench_g3_pairs(torch, dev, pairs=64, steps=10):
    """The reference's SHIPPED pipeline (config_unet_bigdata.py, general_dataset.py:71): UNet(3,1) on the two fingers of a
    320x427 frame pair, area-down-sampled to 160x213, depth up-sampled back to 320x427 -- through
    predict_depth_from_frame_pairs (one gsd_forward per batch: split + difference + resampling + normalisation fused)."""
    import types
    from gelslim_depth_b200.models.unet import UNet
    from gelslim_depth_b200.processing_utils.complete_prediction import predict_depth_from_frame_pairs
    cfg = types.SimpleNamespace(input_tactile_image_size=(160, 213), interp_method="area", norm_scale=0.9,
                                image_normalization_method="0_255_to_0_1", image_normalization_parameters=None,
                                depth_normalization_method="min_max_to_0_-1",
                                depth_normalization_parameters=(-1.9180814027786255, 0.0))
    torch.manual_seed(0)
    net3 = UNet(3, 1, layer_dimensions=DIMS).to(dev).eval()
    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 256, (pairs, 6, H, W), generator=g, dtype=torch.uint8).to(dev)
    base = torch.randint(0, 256, (1, 6, H, W), generator=g, dtype=torch.uint8).float().to(dev)
    for _ in range(3):
        y = predict_depth_from_frame_pairs(frames, base, net3, (H, W), cfg)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        y = predict_depth_from_frame_pairs(frames, base, net3, (H, W), cfg)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    return {"metric": "g3_frame_pairs_per_s", "value": pairs / (ms / 1e3), "ms_per_step": ms, "pairs_per_step": pairs,
            "gflop_per_pair": 99.21, "out_shape": list(y.shape),
            "what": "predict_depth_from_frame_pairs: uint8 6x320x427 pairs -> UNet(3,1) @160x213 on 2 fingers -> 2x320x427 depth (mm)"}


def bench_latency(torch, net, dev, base, frames=500):
    """BASELINE configs[2]: batch-1 streaming through gelslim_depth_b200.streaming.DepthStream -- one interleaved uint8
    camera frame pair in pinned host memory -> H2D -> fused forward -> D2H depth map, one CUDA-graph replay + host
    synchronisation per frame pair; wall-clock per frame pair (host copy into the ring slot included)."""
    import types
    from gelslim_depth_b200.streaming import DepthStream
    cfg = types.SimpleNamespace(input_tactile_image_size=(H, W), interp_method="area", norm_scale=0.9,
                                image_normalization_method="0_255_to_0_1", image_normalization_parameters=None,
                                depth_normalization_method="min_max_to_0_-1",
                                depth_normalization_parameters=(-1.9180814027786255, 0.0))
    ds = DepthStream(net, cfg, (H, W), base_tactile_image=base[0], output_size=(H, W), layout="hwc_u8", frame_pairs=False, slots=4)
    g = torch.Generator().manual_seed(7)
    cam = torch.randint(0, 256, (8, H, W, CIN), generator=g, dtype=torch.uint8)
    for i in range(20):
        ds(cam[i % 8])
    ds.latencies_ms.clear()
    for i in range(frames):
        ds(cam[i % 8])
    p50, p99 = ds.latency_percentiles((0.5, 0.99))
    # zero-copy ingest: the frame is already in the pinned ring slot (a camera driver writes there), as in tools/latency.py
    ds.latencies_ms.clear()
    for i in range(frames):
        ticket, buf = ds.acquire()
        if i < 8:
            buf.copy_(cam[i % 8])
        ds.result(ds.submit(ticket))
    z50, z99 = ds.latency_percentiles((0.5, 0.99))
    return {"metric": "b1_latency_ms_per_frame_pair", "p50_ms": p50, "p99_ms": p99, "frames": frames,
            "zero_copy_ingest": {"p50_ms": z50, "p99_ms": z99, "what": "frame already in the pinned ring slot (acquire/submit)"},
            "what": "DepthStream: uint8 HWC 6x320x427 frame pair -> pinned ring slot -> CUDA-graph replay (H2D + difference image + "
                    "U-Net + de-normalisation + D2H) -> host depth map; wall clock per frame pair incl. host sync",
            "launches_per_replay": ds.plan.launches}


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # libraries that print to fd 1 (e.g. "NCCL version ...") must not pollute the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step (BASELINE configs[1]: 64)")
    ap.add_argument("--chunk", type=int, default=0, help="frames per pipelined chunk of the host path (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-steps", type=int, default=20, help="timed training steps (config 4: bf16, batch 32/GPU); 0 = skip")
    ap.add_argument("--cpu-batch", type=int, default=0, help="--impl reference: frame pairs per CPU step (0 = the faster of 1 and 8)")
    ap.add_argument("--train-batch", type=int, default=32)
    ap.add_argument("--layers", action="store_true", help="print the per-launch table to stderr")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from gelslim_depth_b200.models.unet import UNet
    from gelslim_depth_b200.engine import make_prepost

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    torch.manual_seed(0)
    net = UNet(CIN, NCLS, layer_dimensions=DIMS).to(dev).eval()      # random init (no checkpoints offline)
    raw, base = synthetic_frames(torch, B, seed=rank)
    raw_h = raw.pin_memory()
    x_dev, base_dev = raw.to(dev), base.to(dev)
    y_dev = torch.empty(B, NCLS, H, W, device=dev)
    y_host = torch.empty(B, NCLS, H, W).pin_memory()
    # get_difference_image + '0_255_to_0_1' + 'min_max_to_0_-1' (config_unet_bigdata.py:39-43) fused around the net
    pp = make_prepost(CIN, (H, W), (H, W), use_diff=True, base_batch=1, in_scale=[1 / 255.0], in_shift=[0.0],
                      out_scale=1.9180814027786255 / -0.9, out_shift=-1.9180814027786255)
    plan = net.plan_for(B, H, W, dev)
    packed = net.packed_weights(plan)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident throughput (value)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        plan.forward(x_dev, base_dev, pp, y_dev, packed)
    barrier()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        plan.forward(x_dev, base_dev, pp, y_dev, packed)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    value = world * B * args.steps / (ms / 1e3)

    # ---------------- end to end through the C ABI with HOST buffers (e2e)
    pp8 = make_prepost(CIN, (H, W), (H, W), use_diff=True, base_batch=1, in_scale=[1 / 255.0], in_shift=[0.0],
                       out_scale=1.9180814027786255 / -0.9, out_shift=-1.9180814027786255, input_u8=True)
    raw8_h = raw.to(torch.uint8).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def timed_ms(run):
        """max over ranks of the device time of `run()` (which must leave every result in host memory)"""
        barrier()
        e0.record(stream)
        run()
        e1.record(stream)
        barrier()
        t_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([t_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t)
        return t_ms

    # (1) a stream of batches: gsd_forward_host_async on rotating staging slots -- the upload of step k+1 and the
    #     download of step k-1 overlap the compute of step k; every step still moves its own frames host -> device
    #     and its own depth maps device -> host, and the timed region ends when the last depth map is in host memory
    n_slots = 2
    plan.set_chunk(args.chunk or B)
    y_hosts = [y_host] + [torch.empty(B, NCLS, H, W).pin_memory() for _ in range(n_slots - 1)]
    y_devs = [y_dev] + [torch.empty_like(y_dev) for _ in range(n_slots - 1)]

    def stream_of_batches(src_h, pre, x_devs):
        def run():
            for k in range(e2e_steps):
                s = k % n_slots
                plan.forward_host_async(src_h, base_dev, pre, y_hosts[s], x_devs[s], y_devs[s], packed, slot=s)
            for s in range(n_slots):
                plan.host_wait(s)
        return run

    x_devs = [x_dev] + [torch.empty_like(x_dev) for _ in range(n_slots - 1)]
    run = stream_of_batches(raw_h, pp, x_devs)
    run()
    e2e = world * B * e2e_steps / (timed_ms(run) / 1e3)
    del x_devs
    x8_devs = [torch.empty(raw8_h.shape, dtype=torch.uint8, device=dev) for _ in range(n_slots)]
    run = stream_of_batches(raw8_h, pp8, x8_devs)
    run()
    e2e_u8 = world * B * e2e_steps / (timed_ms(run) / 1e3)
    del y_devs[1:], y_hosts[1:]

    # (2) one blocking call per batch (gsd_forward_host): upload | compute | download pipelined chunk by chunk inside
    #     the call, the first upload and the last download are exposed every step
    chunk = args.chunk or max(1, B // 4)      # 16-frame chunks measured best (profiles/r1_sweep_1gpu.jsonl)
    ramp = max(1, chunk // 2) if not args.chunk else 0   # smaller first / last chunk: less exposed upload / download
    plan.set_chunk(chunk, first=ramp, last=ramp)

    def blocking(src_h, pre, xd):
        def run():
            for _ in range(e2e_steps):
                plan.forward_host(src_h, base_dev, pre, y_host, xd, y_dev, packed)   # returns when y_host is complete
        return run

    run = blocking(raw_h, pp, x_dev)
    run()
    e2e_blk = world * B * e2e_steps / (timed_ms(run) / 1e3)
    run = blocking(raw8_h, pp8, x8_devs[0])
    run()
    e2e_blk_u8 = world * B * e2e_steps / (timed_ms(run) / 1e3)
    del x8_devs
    plan.set_chunk(B)

    # ---------------- training step (BASELINE configs[3]): fwd + MSE + bwd + bucketed all-reduce + Adam + EMA
    train = None
    if args.train_steps > 0:
        del x_dev, y_dev
        net._plans.clear()
        torch.cuda.empty_cache()
        try:
            train = bench_train(torch, dist, dev, rank, world, args.train_batch, args.train_steps)
        except Exception as e:          # noqa: BLE001
            if world > 1:               # the other ranks are inside collectives: fail loudly rather than hang them
                raise
            train = {"error": f"{type(e).__name__}: {e}"}
        net.eval()

    if rank != 0:
        shutdown(dist, world)
        return
    x_dev = raw.to(dev)
    y_dev = torch.empty(B, NCLS, H, W, device=dev)
    plan = net.plan_for(B, H, W, dev)
    packed = net.packed_weights(plan)

    # ---------------- live roofline of the dominant kernels (the 22 conv / transposed-conv GEMM launches) + per-launch table
    # achieved (in-loop)  = algorithmic conv FLOPs of one step / (driver-visible in-loop step time x the convs' share of a
    #                       step), against the SUSTAINED cuBLAS peak: both were measured inside a long back-to-back loop;
    # achieved (isolated) = the same FLOPs / sum of the per-launch CUDA-event times of gsd_forward_profiled (each launch
    #                       bracketed by events: burst conditions), against the BURST peak.
    # FLOPs are the real (unpadded) problem's: inc.0 counts its 6 input channels, not the 16 stored.
    sustained, burst, hbm, src = measured_peaks()
    profs = [plan.forward_profiled(x_dev, base_dev, pp, y_dev, packed) for _ in range(5)]
    prof = [(statistics.median(pr[i][0] for pr in profs), profs[0][i][1]) for i in range(len(profs[0]))]
    conv_ms = sum(m for m, f in prof if f > 0)
    conv_flops = sum(f for m, f in prof if f > 0)
    other_ms = sum(m for m, f in prof if f == 0)
    conv_share = conv_ms / (conv_ms + other_ms)
    step_ms = ms / args.steps
    achieved_loop = conv_flops / (step_ms * conv_share * 1e-3) / 1e12
    achieved_iso = conv_flops / (conv_ms * 1e-3) / 1e12
    first = ["inc.0+prologue(fused)"] if plan.first_fused else ["prologue", "inc.0"]
    names = first + ["inc.3"] + [f"down.{i}.{j}" for i in range(4) for j in (0, 3)] + \
            [f"up.{i}.{n}" for i in range(4) for n in ("up", "conv.0", "conv.3")] + ["head"]
    table = [[n, round(m, 4), round(f / (m * 1e-3) / 1e12, 1) if f else None] for n, (m, f) in zip(names, prof)]
    if args.layers:
        for r in table:
            print(r, file=sys.stderr)
    tr = train or {}
    roofline = {"bound": "tensor", "kernel": "conv_first / conv_halo / conv_tc (22 launches per step: 18 conv3x3 + 4 transposed convs as "
                                             "implicit GEMMs on tcgen05)",
                "achieved": achieved_loop, "peak": sustained, "unit": "TFLOP/s", "frac": achieved_loop / sustained,
                "peak_source": f"{src} bf16_tflops_sustained: `achieved` = conv FLOPs / (in-loop step time x conv share of the step)",
                "isolated": {"achieved": achieved_iso, "peak": burst, "frac": achieved_iso / burst,
                             "what": "same FLOPs / sum of per-launch CUDA-event times (gsd_forward_profiled, median of 5), vs the burst peak"},
                "traffic": (traffic_from_profile() or {}).get("dram_bytes_per_step"),
                "traffic_what": "dram__bytes_read.sum + dram__bytes_write.sum summed over the conv launches of one batch-64 forward "
                                "(committed ncu --set full capture), bytes",
                "traffic_detail": traffic_from_profile(),
                "conv_share_of_step": conv_share, "flops_per_launch_set": conv_flops,
                "algorithmic_gflop_per_frame": GFLOP_PER_FRAME, "prologue_fused_into_first_conv": bool(plan.first_fused),
                # scalars of the secondary legs, mirrored here so that they survive in the driver's parsed record
                "train_samples_per_s": tr.get("value"), "train_ms_per_step": tr.get("ms_per_step"), "train_steps": tr.get("steps"),
                "train_tensor_frac_sustained": tr.get("tensor_frac_sustained"), "ddp_in_sync": tr.get("ddp_in_sync"),
                "ddp_reduced_grad_rel_err": tr.get("ddp_reduced_grad_rel_err")}

    # secondary blocks: a failure there must not take the headline line down with it
    try:
        latency = bench_latency(torch, net, dev, base)
    except Exception as e:          # noqa: BLE001
        latency = {"error": f"{type(e).__name__}: {e}"}
    try:
        g3 = bench_g3_pairs(torch, dev)
    except Exception as e:          # noqa: BLE001
        g3 = {"error": f"{type(e).__name__}: {e}"}
    roofline["latency_b1_p50_ms"], roofline["latency_b1_p99_ms"] = latency.get("p50_ms"), latency.get("p99_ms")

    cpu = None
    if not args.no_cpu_baseline:
        fps, cores, sample = cpu_reference_fps(torch)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}

    # bulky tables first, scalars last: the driver keeps the tail of the line
    line = {"layers": table, "layers_columns": ["launch", "ms (isolated, median of 5)", "TFLOP/s"],
            "train_losses": tr.pop("losses", None) if isinstance(tr, dict) else None,
            "g3_pipeline": g3, "latency": latency, "train": train,
            "metric": "unet_frames_per_s_6x320x427", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(B),
            "tensor_frac_whole_step": value / world * GFLOP_PER_FRAME / 1e3 / sustained,
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_u8, "unit": "frames/s", "h2d_bytes_per_step": B * CIN * H * W,
                    "d2h_bytes_per_step": B * NCLS * H * W * 4, "steps": e2e_steps,
                    "api": "gsd_forward_host_async + gsd_forward_host_wait: a stream of batches on 2 rotating staging slots; pinned "
                           "uint8 camera frames in (gsd_prepost.input_u8, what README.md:155-171's capture loop delivers), fp32 "
                           "depth maps out; every step's upload and download are inside the timed region, which ends when the "
                           "last depth map is in host memory",
                    "slots": n_slots,
                    "fp32_frames": {"value": e2e, "h2d_bytes_per_step": B * CIN * H * W * 4,
                                    "what": "same with the frames already converted to fp32 on the host (the reference's tensor type)"},
                    "blocking_call": {"value": e2e_blk_u8, "fp32_frames": e2e_blk, "chunk_frames": chunk,
                                      "first_last_chunk_frames": ramp,
                                      "what": "one gsd_forward_host call per batch, each returning only when its depth "
                                              "maps are in host memory (first upload / last download exposed)"}},
            "gpu_launches": plan.launches * args.steps, "clocks": clocks,
            "summary": {"frames_per_s": round(value, 1), "e2e_frames_per_s": round(e2e_u8, 1), "ms_per_step": round(step_ms, 3),
                        "roofline_frac_sustained": round(achieved_loop / sustained, 4),
                        "roofline_frac_isolated_burst": round(achieved_iso / burst, 4),
                        "train_samples_per_s": tr.get("value"), "train_ms_per_step": tr.get("ms_per_step"),
                        "train_steps": tr.get("steps"), "train_tensor_frac_sustained": tr.get("tensor_frac_sustained"),
                        "ddp_in_sync": tr.get("ddp_in_sync"), "ddp_reduced_grad_rel_err": tr.get("ddp_reduced_grad_rel_err"),
                        "latency_b1_p50_ms": latency.get("p50_ms"), "latency_b1_p99_ms": latency.get("p99_ms"),
                        "g3_pairs_per_s": g3.get("value"), "cpu_frames_per_s": cpu["value"] if cpu else None}}
    emit(line)
    shutdown(dist, world)


if __name__ == "__main__":
    main()
