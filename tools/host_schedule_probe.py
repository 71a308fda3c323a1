"""Chunk schedules of the host-buffer call (BCAD_HOST_SIZES), 512 images: float32 in/out and 8-bit in/out.

    python tools/host_schedule_probe.py "64,128,192,128" "128,128,128,128" ...
"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bcad_b200
import bench

B = 512
spec = bcad_b200.NetSpec.torch_flavour((256, 256, 1), 2, [(32, 3), (64, 3)], [256, 128], 0.01)
cw, cb, dw, db = bench.synth_weights()
eng = bcad_b200.Engine(spec, precision="fp16", max_batch=B)
eng.set_weights(cw, cb, dw, db)
x = torch.from_numpy(np.concatenate([bench.synth_images(64, (256, 256, 1), seed=1)] * 8)).pin_memory()
x8 = torch.from_numpy(np.clip(np.rint(x.numpy() * 255.0), 0, 255).astype(np.uint8)).pin_memory()
hf = torch.empty((B, 256, 256), dtype=torch.float32).pin_memory()
h8 = torch.empty((B, 256, 256), dtype=torch.uint8).pin_memory()


def run(fn, n=20):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


for sched in sys.argv[1:] or ["default"]:
    if sched == "default":
        os.environ.pop("BCAD_HOST_SIZES", None)
    else:
        os.environ["BCAD_HOST_SIZES"] = sched
    f = run(lambda: eng.predict_explain_host(x.numpy(), None, "logit", heat_out=hf.numpy()))
    u = run(lambda: eng.predict_explain_host(x8.numpy(), None, "logit", heat_out=h8.numpy(), heat_dtype=np.uint8))
    print(f"{sched:>28}: float32 {f:.3f} ms   u8 in/out {u:.3f} ms", flush=True)
