"""The oracle against hand-derived known answers (tests/kat_cases.py) -- CPU."""
import numpy as np

import kat_cases as K
from util import ocnn


def _cfg(case, flavour, alpha=0.01):
    mk = ocnn.NetConfig.numpy_flavour if flavour == "numpy" else ocnn.NetConfig.torch_flavour
    cfg = mk(case["input_shape"], 2, case["conv"], case["hidden"], alpha)
    if flavour == "torch":
        cfg = ocnn.NetConfig(cfg.input_shape, 2, cfg.conv_layers, cfg.hidden_units, alpha, alpha, 0, cfg.flatten, cfg.pool_ties, cfg.head)
    return cfg, ocnn.Params(case["conv_w"], case["conv_b"], case["dense_w"], case["dense_b"])


def test_impulse_conv_orientation_and_layout():
    c = K.impulse_conv()
    cfg, p = _cfg(c, "numpy")
    np.testing.assert_allclose(ocnn.forward(cfg, p, c["x"]).conv_out[0].numpy(), c["conv_out"], atol=1e-12)


def test_pool_tie_rules():
    c = K.pool_tie()
    for flavour, want in (("numpy", c["dA_dup"]), ("torch", c["dA_first"])):
        cfg, p = _cfg(c, flavour)
        cache = ocnn.forward(cfg, p, c["x"])
        np.testing.assert_allclose(cache.pool_out[0].numpy().reshape(1, 2), c["pooled"], atol=0)
        cag, _, _ = ocnn.backward(cfg, p, cache, ocnn.top_gradient(cache, np.array([0]), "logit"), through_input=False)
        np.testing.assert_allclose(cag[0].numpy(), want, atol=1e-12)


def test_flatten_order():
    c = K.flatten_order()
    for flavour, want in (("numpy", c["logit0_hwc"]), ("torch", c["logit0_chw"])):
        cfg, p = _cfg(c, flavour)
        assert abs(float(ocnn.forward(cfg, p, c["x"]).logits[0, 0]) - want) < 1e-12


def test_softmax_clip():
    c = K.softmax_clip()
    cfg, p = _cfg(c, "numpy")
    assert abs(float(ocnn.forward(cfg, p, c["x"]).probs[0, 1]) / c["p1_clip"] - 1) < 1e-9
    cfg, p = _cfg(c, "torch")
    assert abs(float(ocnn.forward(cfg, p, c["x"]).probs[0, 1]) / c["p1_plain"] - 1) < 1e-9


def test_leaky_relu_at_exactly_zero():
    c = K.leaky_at_zero()
    cfg, p = _cfg(c, "numpy", alpha=c["alpha"])
    cache = ocnn.forward(cfg, p, c["x"])
    np.testing.assert_allclose(cache.logits.numpy(), c["logits"], atol=1e-12)
    cag, _, _ = ocnn.backward(cfg, p, cache, ocnn.top_gradient(cache, np.array([0]), "logit"), through_input=False)
    # every element of a constant window ties: the duplicated gradient is g_pool at all four positions
    want = np.repeat(np.repeat(c["g_pool"].reshape(1, 1, 2, 1), 2, axis=1), 2, axis=2)
    np.testing.assert_allclose(cag[0].numpy(), want, atol=1e-12)


def test_every_image_comparison_helper_with_a_float32_stand_in_engine():
    """oracle/compare.py (the helper behind the 2048-image GPU gate and bench.py's `check`) on the CPU: an engine stand-in that
    evaluates the oracle in float32 passes at float32-grade bounds on every image; one that perturbs a hidden pre-activation
    across 0 by more than tau is reported as a mask violation, and a flipped class as a class mismatch."""
    import torch
    from oracle import cnn as ocnn, gradcam as ogc
    from oracle.compare import compare_all_images
    cfg = ocnn.NetConfig.torch_flavour((24, 24, 1), 2, [(4, 3), (8, 3)], [12, 6], 0.01)
    p = ocnn.init_params(cfg, seed=5, bias_std=0.05)
    x = ocnn.synth_images(9, (24, 24, 1), seed=3)

    class Stand:
        max_batch = 4

        def __init__(self, bump=0.0, flip=False):
            self.bump, self.flip = bump, flip

        def predict_explain(self, xs, class_idx, mode):
            c = ocnn.forward(cfg, p, xs, dtype=torch.float32)
            if self.bump:
                z = c.z[0].clone()
                z[0, 0] = -z[0, 0] + (self.bump if z[0, 0] < 0 else -self.bump)      # unit 0 of image 0 lands on the other side
                c.z[0] = z
            cls = c.logits.argmax(dim=-1)
            if self.flip:
                cls = 1 - cls
            ci = cls.numpy() if class_idx is None else class_idx
            cag, _, _ = ocnn.backward(cfg, p, c, ocnn.top_gradient(c, ci, mode), through_input=False)
            heat = ogc.gradcam_tail_nhwc(c.conv_out[1].numpy(), cag[1].numpy(), (24, 24))
            self._z = [z.clone() for z in c.z]
            return cls.int(), c.probs, c.logits, torch.from_numpy(heat)

        def get_tensor(self, kind, j, B):
            return self._z[j]

    modes = [(None, "logit"), (np.arange(9) % 2, "softmax_ce")]
    for r in compare_all_images(cfg, p, x, Stand(), modes, tau=1e-5):
        assert r["mask_violations"] == 0 and r["cls_equal"].all() and not r["overridden"].any()
        assert r["logit_err"].max() < 1e-5 and r["heat_err"].max() < 1e-4
    r = compare_all_images(cfg, p, x, Stand(bump=0.5), modes, tau=1e-5)[0]
    assert r["mask_violations"] >= 1 and r["overridden"][0]
    r = compare_all_images(cfg, p, x, Stand(flip=True), modes, tau=1e-5)[0]
    assert not r["cls_equal"].any()
