"""Diagnostic (GPU box, any number of GPUs): what the HOST side can move when N GPUs copy at once -- the ceiling of the e2e leg.

One thread per GPU, each with its own pinned input / output buffers of the bench step's size (134 MB each way) and two streams
(H2D and D2H concurrently, no compute); all threads start together.  Prints the aggregate GB/s for N = 1, 2, 4, 8 (as available)
and what that means in images/s for the float32 call (268 MB per 512 images) and the uint8 call (67 MB).
    python tools/host_ceiling_probe.py [--mb 128] [--iters 10] [--numa]      (--numa: bind each thread to its GPU's NVML CPU set)
"""
import argparse
import json
import os
import threading
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=128)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--numa", action="store_true")
args = ap.parse_args()
ngpu = torch.cuda.device_count()
n = args.mb * 1024 * 1024 // 4
bufs = []
for d in range(ngpu):
    dev = torch.device("cuda", d)
    bufs.append((torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory(),
                 torch.empty(n, dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.float32, device=dev),
                 torch.cuda.Stream(dev), torch.cuda.Stream(dev)))


def bind(d):
    try:
        import pynvml
        pynvml.nvmlInit()
        mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(d), (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:
        print("bind skipped:", e)


def worker(d, go, iters, times):
    if args.numa:
        bind(d)
    h_in, h_out, d_in, d_out, s1, s2 = bufs[d]
    dev = torch.device("cuda", d)
    torch.cuda.set_device(dev)
    go.wait()
    t0 = time.perf_counter()
    for _ in range(iters):
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(dev)
    times[d] = time.perf_counter() - t0


out = {}
for N in (1, 2, 4, 8):
    if N > ngpu:
        break
    for rep in range(2):                     # first pass warms the streams up
        go = threading.Barrier(N)
        times = {}
        ths = [threading.Thread(target=worker, args=(d, go, args.iters, times)) for d in range(N)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    dt = max(times.values())
    gbs = N * 2 * args.mb / 1024 * args.iters / dt
    out[N] = {"aggregate_GBps": gbs, "f32_call_images_per_s": gbs * 1e9 / (268.4e6 / 512), "u8_call_images_per_s": gbs * 1e9 / (67.1e6 / 512)}
    print(f"N={N}: {gbs:.1f} GB/s aggregate (H2D + D2H at once) -> float32 call <= {out[N]['f32_call_images_per_s'] / 1e3:.0f} k images/s, "
          f"uint8 call <= {out[N]['u8_call_images_per_s'] / 1e3:.0f} k images/s")
print(json.dumps(out))
