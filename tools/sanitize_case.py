"""Small end-to-end case for compute-sanitizer (memcheck): both tensor modes and the fp32 path once."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bcad_b200
from oracle import cnn as ocnn
cfg = ocnn.NetConfig.torch_flavour((64, 64, 1), 2, [(32, 3), (64, 3)], [64, 32], 0.01)
p = ocnn.init_params(cfg, seed=7, bias_std=0.05)
x = ocnn.synth_images(5, (64, 64, 1), seed=1)
spec = bcad_b200.NetSpec.torch_flavour((64, 64, 1), 2, [(32, 3), (64, 3)], [64, 32], 0.01)
for prec in ("fp16", "fp16x3", "fp32"):
    e = bcad_b200.Engine(spec, precision=prec, max_batch=4)
    e.set_weights(p.conv_w, p.conv_b, p.dense_w, p.dense_b)
    cls, probs, logits, heat = e.predict_explain(x, None, "logit")
    h = e.predict_explain_host(x, None, "logit")
    x8 = np.clip(np.rint(x * 255.0), 0, 255).astype(np.uint8)
    h8 = e.predict_explain_host(x8, None, "logit", heat_dtype=np.uint8)       # 8-bit pixels in, heatmap_uint8 out
    print(prec, cls.tolist(), float(heat.max()), h[0].tolist(), int(h8[3].max()))
    e.close()
    if prec == "fp16":                                                       # the fused two-block kernel on a small case
        os.environ["BCAD_FUSED_CONV"] = "1"
        e = bcad_b200.Engine(spec, precision=prec, max_batch=4)
        e.set_weights(p.conv_w, p.conv_b, p.dense_w, p.dense_b)
        print("fused", e.predict_explain(x, None, "logit")[0].tolist())
        e.close()
        del os.environ["BCAD_FUSED_CONV"]
print("done")
