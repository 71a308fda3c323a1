"""Shared helpers for the tests: oracle config/params <-> product Engine."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cnn as ocnn  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def spec_from_cfg(cfg):
    import bcad_b200
    return bcad_b200.NetSpec(tuple(cfg.input_shape), cfg.num_classes, list(cfg.conv_layers), list(cfg.hidden_units),
                             cfg.alpha_conv, cfg.alpha_dense, cfg.pad, cfg.flatten, cfg.pool_ties, cfg.head)


def engine_from(cfg, params, **kw):
    import bcad_b200
    eng = bcad_b200.Engine(spec_from_cfg(cfg), **kw)
    eng.set_weights(params.conv_w, params.conv_b, params.dense_w, params.dense_b)
    return eng


def load_numpy_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = ocnn.NetConfig.numpy_flavour(tuple(int(v) for v in g["input_shape"]), 2,
                                       [tuple(int(a) for a in r) for r in g["conv_layers"]],
                                       [int(v) for v in g["hidden"]], float(g["alpha"]))
    n_conv = len(cfg.conv_layers)
    conv_idx = [2 * i for i in range(n_conv)]
    dense_idx = [2 * n_conv + j for j in range(len(cfg.hidden_units) + 1)]
    p = ocnn.Params([g[f"W{i}"] for i in conv_idx], [g[f"b{i}"] for i in conv_idx],
                    [g[f"W{i}"] for i in dense_idx], [g[f"b{i}"] for i in dense_idx])
    return g, cfg, p, conv_idx, dense_idx


def load_torch_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = ocnn.NetConfig.torch_flavour(tuple(int(v) for v in g["input_shape"]), 2,
                                       [tuple(int(a) for a in r) for r in g["conv_layers"]],
                                       [int(v) for v in g["hidden"]], float(g["alpha"]))
    n_conv, n_dense = len(cfg.conv_layers), len(cfg.hidden_units) + 1
    p = ocnn.Params([g[f"sd.convs.{i}.weight"].transpose(0, 2, 3, 1) for i in range(n_conv)],
                    [g[f"sd.convs.{i}.bias"] for i in range(n_conv)],
                    [g[f"sd.fc.{3 * j}.weight"] for j in range(n_dense)],
                    [g[f"sd.fc.{3 * j}.bias"] for j in range(n_dense)])
    return g, cfg, p


def _switches_from(A_nhwc: np.ndarray, ties: str) -> np.ndarray:
    """Pool switches (float mask) of an activation map, per the flavour's tie rule."""
    import torch
    _, sw = ocnn._pool_with_switches(torch.as_tensor(A_nhwc).permute(0, 3, 1, 2), ties)
    return sw.permute(0, 2, 3, 1).numpy()


def near_tie_windows(A_nhwc: np.ndarray, rel=1e-6) -> np.ndarray:
    """[B,h2,w2,C] bool: pool windows holding a value within `rel` (relative) of the maximum without
    being equal to it -- a float32 computation may legitimately see an exact tie there (or break one)."""
    b, h, w, c = A_nhwc.shape
    h2, w2 = h // 2, w // 2
    win = A_nhwc[:, :2 * h2, :2 * w2].reshape(b, h2, 2, w2, 2, c).transpose(0, 1, 3, 5, 2, 4).reshape(b, h2, w2, c, 4)
    mx = win.max(axis=-1, keepdims=True)
    gap = mx - win
    return ((gap > 0) & (gap < rel * np.maximum(np.abs(mx), 1e-30))).any(axis=-1)


def oracle_heatmaps(cfg, params, x, class_idx, grad_mode, A_for_ties=None):
    """Oracle predict + Grad-CAM at the last conv block (float64 net, float32 tail as pytorch_grad_cam).

    A_for_ties: optional [B,h,w,F] activation map whose EXACT-EQUALITY structure decides the pool switches of
    the last block (the tie rule compares floats for equality, so it can only be reproduced bit-for-bit at the
    precision the activations were computed in; see DESIGN.md "tie semantics")."""
    import torch
    from oracle import gradcam as ogc
    cache = ocnn.forward(cfg, params, x)
    score = cache.probs if cfg.head == "softmax" else cache.logits
    cls = score.argmax(dim=-1).numpy()
    ci = cls if class_idx is None else class_idx
    if A_for_ties is not None:
        cache.switches[-1] = torch.as_tensor(_switches_from(A_for_ties, cfg.pool_ties)).to(cache.logits.dtype)
    cag, _, _ = ocnn.backward(cfg, params, cache, ocnn.top_gradient(cache, ci, grad_mode), through_input=False)
    last = len(cfg.conv_layers) - 1
    A = cache.conv_out[last].numpy().astype(np.float32)
    dA = cag[last].numpy().astype(np.float32)
    heat = ogc.gradcam_tail_nhwc(A, dA, cfg.input_shape[:2])
    return cls, cache, A, dA, heat


from oracle.compare import compare_all_images  # noqa: E402,F401  (shared with bench.py's untimed `check` leg)
