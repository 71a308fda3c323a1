"""Host-side cost of each call of one training step (sync after every phase): where does the CPU time go?"""
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
    ocnn, cfg, params = bench.oracle_setup()
    B = 64
    spec = bcad_b200.NetSpec.torch_flavour(bench.INPUT_SHAPE, 2, bench.CONV_LAYERS, bench.HIDDEN, 0.01)
    eng = bcad_b200.Engine(spec, precision="fp32", max_batch=B, keep_all_activations=True)
    eng.set_weights(params.conv_w, params.conv_b, params.dense_w, params.dense_b)
    x_host = torch.from_numpy(ocnn.synth_images(B, bench.INPUT_SHAPE, seed=1)).pin_memory()
    y = torch.from_numpy((np.arange(B) % 2).astype(np.int32)).cuda()
    grads = None
    for it in range(6):
        t = [time.perf_counter()]
        def lap(sync=True):
            if sync:
                torch.cuda.synchronize()
            t.append(time.perf_counter())
        x = x_host.cuda(non_blocking=True); lap(False); lap()
        eng.predict(x); lap(False); lap()
        grads, loss = eng.train_backward(x, y, grads); lap(False); lap()
        eng.apply_update(grads, "adam", lr=1e-4); lap(False); lap()
        v = float(loss.mean()); lap()
        d = [(t[i + 1] - t[i]) * 1e3 for i in range(len(t) - 1)]
        print(f"it{it}: h2d call {d[0]:.2f} wait {d[1]:.2f} | predict call {d[2]:.2f} wait {d[3]:.2f} | backward call {d[4]:.2f} wait {d[5]:.2f} | "
              f"update call {d[6]:.2f} wait {d[7]:.2f} | loss {d[8]:.2f}  total {sum(d):.2f} ms")


if __name__ == "__main__":
    main()
