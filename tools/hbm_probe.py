"""Diagnostic (GPU box): pure-write vs copy HBM bandwidth (torch fill_/copy_ on 2 GiB), CUDA events."""
import torch
n = 1 << 29   # 2 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
w = t(lambda: a.fill_(1.0))
c = t(lambda: b.copy_(a))
r = t(lambda: a.sum())
print(f"write-only {n*4/w/1e9:.0f} GB/s ; copy (read+write bytes) {2*n*4/c/1e9:.0f} GB/s ; read-only (sum) {n*4/r/1e9:.0f} GB/s")
