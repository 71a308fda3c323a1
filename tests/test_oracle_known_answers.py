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
