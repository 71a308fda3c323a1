"""Tiny U-Net encoder front on the GPU vs the reference-generated fixtures and the oracle."""
import os

import numpy as np
import pytest

from util import GOLDEN

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.mark.parametrize("name,c", [("ref_unet_small", 1), ("ref_unet_odd", 2)])
def test_unet_front_vs_reference_fixture(name, c):
    from bcad_b200 import unet as U
    from oracle import unet as ou
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    ks = ou.draw_kernels(c, int(g["seed"]))
    c1 = U.conv2d(g["x"], ks[0])
    assert c1.shape == g["c1"].shape
    assert np.all(c1[:, -2:] == 0) and np.all(c1[:, :, -2:] == 0)
    assert _rel(c1, g["c1"]) <= 1e-5
    assert _rel(U.max_pool(U.relu(g["c1"])), g["p1"]) <= 1e-6
    bn = U.tiny_unet(g["x"], ks)
    assert bn.shape == g["bn"].shape and _rel(bn, g["bn"]) <= 1e-4       # randn(std 1) kernels: values in the 100s
    assert _rel(U.average_pool(g["bn"], 3), g["avg3"]) <= 1e-6
    np.random.seed(int(g["seed"]))                                         # the reference's own entry point + RNG stream
    assert _rel(U.tiny_unet_numpy(g["x"]), g["bn"]) <= 1e-4


def test_unet_front_feeds_the_cnn_full_size():
    """BASELINE config 3 shape flow: 256x256x1 -> tiny U-Net (67x67x64) -> average_pool(3) (22x22x64) -> CNN + Grad-CAM."""
    import torch
    from bcad_b200 import unet as U
    from oracle import unet as ou
    from oracle import cnn as ocnn
    from util import engine_from, oracle_heatmaps
    x = ocnn.synth_images(3, (256, 256, 1), seed=5)
    ks = [k * s for k, s in zip(ou.draw_kernels(1, 7), (0.3, 0.08, 0.06))]        # scaled so activations stay O(1)
    bn = U.tiny_unet(x, ks)
    want = ou.tiny_unet(x.astype(np.float64), ks)
    assert bn.shape == (3, 67, 67, 64) and _rel(bn, want) <= 1e-4
    feat = U.average_pool(bn, 3)
    assert feat.shape == (3, 22, 22, 64)
    cfg = ocnn.NetConfig.torch_flavour((22, 22, 64), 2, [(32, 3), (64, 3)], [64, 32], 0.01)
    p = ocnn.init_params(cfg, seed=7, bias_std=0.05)
    eng = engine_from(cfg, p, max_batch=4)
    cls, probs, logits, heat = eng.predict_explain(feat.astype(np.float32), None, "logit")
    o_cls, cache, A, dA, o_heat = oracle_heatmaps(cfg, p, feat.astype(np.float32), None, "logit")
    assert np.array_equal(cls.cpu().numpy(), o_cls)
    assert np.abs(logits.cpu().numpy() - cache.logits.numpy()).max() <= 1e-4 * max(1.0, np.abs(cache.logits.numpy()).max())
    assert np.abs(heat.cpu().numpy() - o_heat).max() <= 1e-4
    eng.close()


def test_process_bottleneck_features_matches_reference():
    """app.py:466-489 outputs recorded from the reference (tests/golden/ref_bottleneck.npz): bit-exact."""
    import os
    import numpy as np
    import torch
    from util import GOLDEN
    from bcad_b200 import bottleneck as bn
    g = np.load(os.path.join(GOLDEN, "ref_bottleneck.npz"))
    got0 = bn.process_bottleneck_features(torch.from_numpy(g["feat0"]), resize_shape=(8, 8))        # tensor: [C,H,W]
    got1 = bn.process_bottleneck_features(g["feat1"], resize_shape=(11, 7))                         # ndarray CHW, odd sizes
    got2 = bn.process_bottleneck_features(g["feat2"], resize_shape=(9, 10))                         # ndarray already HWC
    for got, want in ((got0, g["out0"]), (got1, g["out1"]), (got2, g["out2"])):
        assert got.shape == want.shape and got.dtype == np.float32
        assert np.array_equal(got, want), float(np.abs(got - want).max())
    # <= 4 channels take OpenCV's double-coordinate path (the Grad-CAM map's): checked against the oracle
    from oracle import gradcam as ogc
    rng = np.random.default_rng(5)
    f = rng.standard_normal((3, 41, 29)).astype(np.float32)
    want = ogc.process_bottleneck_features(f, (13, 17))
    assert np.abs(bn.process_bottleneck_features(f, (13, 17)) - want).max() <= 1e-6
    # batched entry point, 64 channels, 256 -> 32 like the app
    fb = rng.standard_normal((2, 64, 256, 256)).astype(np.float32)
    out = bn.resize_batch(fb, (32, 32)).cpu().numpy()
    assert np.array_equal(out[1], ogc.process_bottleneck_features(fb[1], (32, 32)))
