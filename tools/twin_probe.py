"""Diagnostic (GPU): per-kernel times of a small fp16x3 handle (the refinement twin's shape: max_batch 16) for 1, 2 and 16 images."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bcad_b200
import bench
spec = bcad_b200.NetSpec.torch_flavour((256, 256, 1), 2, [(32, 3), (64, 3)], [256, 128], 0.01)
eng = bcad_b200.Engine(spec, precision="fp16x3", max_batch=16)
eng.set_weights(*bench.synth_weights())
x = torch.from_numpy(bench.synth_images(16, (256, 256, 1), seed=1)).cuda()
eng.set_profiling(True)
for n in (1, 2, 16):
    for _ in range(5):
        eng.predict_explain(x[:n], None, "logit")
    torch.cuda.synchronize()
    acc = {}
    for _ in range(5):
        eng.predict_explain(x[:n], None, "logit")
        torch.cuda.synchronize()
        for name, ms in eng.last_profile():
            acc[name] = acc.get(name, 0.0) + ms / 5
    print(f"n={n}: total {sum(acc.values()):.4f} ms | " + " | ".join(f"{k} {v:.4f}" for k, v in acc.items()))
