"""Diagnostic (GPU box): pinned H2D / D2H bandwidth, alone and concurrently, for several transfer sizes."""
import time
import torch
dev = torch.device("cuda", 0)
for mb in (4, 16, 32, 128):
    n = mb * 1024 * 1024 // 4
    h_in = torch.empty(n, dtype=torch.float32).pin_memory()
    h_out = torch.empty(n, dtype=torch.float32).pin_memory()
    d_in = torch.empty(n, dtype=torch.float32, device=dev)
    d_out = torch.empty(n, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(h2d, d2h, reps=10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    run(True, True, 2)
    a, b, c = run(True, False), run(False, True), run(True, True)
    print(f"{mb} MB: H2D {mb/1024/a:.1f} GB/s  D2H {mb/1024/b:.1f} GB/s  both: {2*mb/1024/c:.1f} GB/s aggregate ({c*1e3:.2f} ms)")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout)
import os
print("cpus", os.cpu_count())
