"""Diagnostic (GPU): 16-bit vs fp32 path on the bench workload -- heat-map error against the cam conditioning
kappa = max_p sqrt(sum_k (alpha_k A_k(p))^2) / (cam_max - cam_min), and the class-margin distribution."""
import sys, os, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bcad_b200
from bcad_b200 import _lib
from oracle import cnn as ocnn

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
FAST = sys.argv[2] if len(sys.argv) > 2 else "fp16"
shape = (256, 256, 1)
cfg = ocnn.NetConfig.torch_flavour(shape, 2, [(32, 3), (64, 3)], [256, 128], 0.01)
p = ocnn.init_params(cfg, seed=7)
x = ocnn.synth_images(N, shape, seed=20251018)
spec = bcad_b200.NetSpec.torch_flavour(shape, 2, [(32, 3), (64, 3)], [256, 128], 0.01)
res = {}
for prec in ("fp32", FAST):
    e = bcad_b200.Engine(spec, precision=prec, max_batch=N)
    e.set_weights(p.conv_w, p.conv_b, p.dense_w, p.dense_b)
    cls, probs, logits, heat = e.predict_explain(x, None, "logit")
    res[prec] = (cls.cpu().numpy(), logits.cpu().numpy(), heat.cpu().numpy())
    if prec == "fp32":
        A = e.get_tensor(_lib.T_CONV_OUT, 1, N).reshape(N, 128, 128, 64)
        al = e.get_tensor(_lib.T_ALPHA, 0, N)
        terms = A * al[:, None, None, :]
        cam = terms.sum(-1).clamp_min(0)
        nrm = terms.pow(2).sum(-1).sqrt().amax(dim=(1, 2))
        rng_ = cam.amax(dim=(1, 2)) - cam.amin(dim=(1, 2))
        kappa = (nrm / rng_.clamp_min(1e-30)).cpu().numpy()
    e.close()
err = np.abs(res[FAST][2] - res["fp32"][2]).reshape(N, -1).max(1)
margin = np.abs(res["fp32"][1][:, 0] - res["fp32"][1][:, 1])
lerr = np.abs(res[FAST][1] - res["fp32"][1]).max(1)
flips = int((res[FAST][0] != res["fp32"][0]).sum())
order = np.argsort(-err)
print("N", N, "class flips", flips, "min margin", margin.min(), "max logit err", lerr.max())
print("heat err: median %.4f p90 %.4f p99 %.4f max %.4f ; >1e-2: %d" % (np.median(err), np.quantile(err, .9), np.quantile(err, .99), err.max(), (err > 1e-2).sum()))
print("kappa: median %.1f p90 %.1f max %.1f" % (np.median(kappa), np.quantile(kappa, .9), kappa.max()))
for i in order[:12]:
    print("img %3d err %.4f kappa %8.1f err/kappa %.2e margin %.4f" % (i, err[i], kappa[i], err[i] / kappa[i], margin[i]))
print("corr(log err, log kappa)", np.corrcoef(np.log(err + 1e-9), np.log(kappa))[0, 1])
