"""Single-request latency of the host-buffer call (the Flask app's flow is one image per request): ms per call at small batches."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bcad_b200  # noqa: E402
import bench  # noqa: E402


def main():
    spec = bcad_b200.NetSpec.torch_flavour(bench.INPUT_SHAPE, 2, bench.CONV_LAYERS, bench.HIDDEN, 0.01)
    w = bench.synth_weights()
    for precision in ("fp16", "fp32"):
        eng = bcad_b200.Engine(spec, precision=precision, max_batch=64)
        eng.set_weights(*w)
        for B in (1, 2, 8, 32):
            x = torch.from_numpy(bench.synth_images(B, bench.INPUT_SHAPE, seed=3)).pin_memory().numpy()
            heat = torch.empty((B, 256, 256), dtype=torch.float32).pin_memory().numpy()
            for _ in range(5):
                eng.predict_explain_host(x, None, "logit", heat_out=heat)
            t0 = time.perf_counter()
            n = 50
            for _ in range(n):
                eng.predict_explain_host(x, None, "logit", heat_out=heat)
            dt = (time.perf_counter() - t0) / n
            print(f"{precision} B={B}: {dt * 1e3:.3f} ms per call, {B / dt:.0f} images/s")
        eng.close()


if __name__ == "__main__":
    main()
