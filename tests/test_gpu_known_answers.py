"""The CUDA path against hand-derived known answers (tests/kat_cases.py) and shape-generic property tests (hypothesis)."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, example, given, settings, strategies as st

import kat_cases as K
from util import engine_from, oracle_heatmaps, ocnn

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def o_heat_raw(cfg, p, cache, classes, mode):
    """The un-normalised low-resolution cam of the oracle (its value range conditions the min-max normalisation)."""
    cag, _, _ = ocnn.backward(cfg, p, cache, ocnn.top_gradient(cache, classes, mode), through_input=False)
    last = len(cfg.conv_layers) - 1
    A, dA = cache.conv_out[last].numpy(), cag[last].numpy()
    alpha = dA.mean(axis=(1, 2), keepdims=True)
    return np.maximum((alpha * A).sum(axis=-1), 0.0)


def _cfg(case, flavour, alpha=0.01):
    mk = ocnn.NetConfig.numpy_flavour if flavour == "numpy" else ocnn.NetConfig.torch_flavour
    cfg = mk(case["input_shape"], 2, case["conv"], case["hidden"], alpha)
    if flavour == "torch":                                   # valid conv keeps the hand-derived shapes; the other torch switches stay
        cfg = ocnn.NetConfig(cfg.input_shape, 2, cfg.conv_layers, cfg.hidden_units, alpha, alpha, 0, cfg.flatten, cfg.pool_ties, cfg.head)
    return cfg, ocnn.Params(case["conv_w"], case["conv_b"], case["dense_w"], case["dense_b"])


def test_impulse_conv_orientation_and_layout():
    from bcad_b200 import _lib
    c = K.impulse_conv()
    cfg, p = _cfg(c, "numpy")
    eng = engine_from(cfg, p, max_batch=1, keep_all_activations=True)
    eng.predict(c["x"])
    got = _np(eng.get_tensor(_lib.T_CONV_OUT, 0, 1)).reshape(c["conv_out"].shape)
    np.testing.assert_allclose(got, c["conv_out"], atol=1e-6)
    eng.close()


def test_pool_tie_rules():
    c = K.pool_tie()
    for flavour, want in (("numpy", c["dA_dup"]), ("torch", c["dA_first"])):
        cfg, p = _cfg(c, flavour)
        eng = engine_from(cfg, p, max_batch=1, keep_all_activations=True)
        eng.predict(c["x"])
        outs, _ = eng.explain_backward(1, np.array([0]), "logit", want_conv=(0,))
        np.testing.assert_allclose(_np(outs[0]), want, atol=1e-6)
        eng.close()


def test_flatten_order():
    c = K.flatten_order()
    for flavour, want in (("numpy", c["logit0_hwc"]), ("torch", c["logit0_chw"])):
        cfg, p = _cfg(c, flavour)
        eng = engine_from(cfg, p, max_batch=1)
        _, _, logits = eng.predict(c["x"])
        assert abs(float(logits[0, 0]) - want) < 1e-4
        eng.close()


def test_softmax_clip():
    c = K.softmax_clip()
    for flavour, want in (("numpy", c["p1_clip"]), ("torch", c["p1_plain"])):
        cfg, p = _cfg(c, flavour)
        eng = engine_from(cfg, p, max_batch=1)
        _, probs, _ = eng.predict(c["x"])
        assert abs(float(probs[0, 1]) / want - 1) < 1e-5
        eng.close()


def test_leaky_relu_at_exactly_zero():
    c = K.leaky_at_zero()
    cfg, p = _cfg(c, "numpy", alpha=c["alpha"])
    eng = engine_from(cfg, p, max_batch=1, keep_all_activations=True)
    _, _, logits = eng.predict(c["x"])
    np.testing.assert_allclose(_np(logits), c["logits"], atol=1e-6)
    outs, _ = eng.explain_backward(1, np.array([0]), "logit", want_conv=(0,))
    want = np.repeat(np.repeat(c["g_pool"].reshape(1, 1, 2, 1), 2, axis=1), 2, axis=2)
    np.testing.assert_allclose(_np(outs[0]), want, atol=1e-6)
    eng.close()


# ------------------------------------------------------------------------------------------------------------------
# property tests: any (H, W, C), odd sizes (floor pooling), kernel sizes 1..3, 1-2 conv blocks, 0-2 hidden layers, both flavours
# ------------------------------------------------------------------------------------------------------------------
@st.composite
def nets(draw):
    flavour = draw(st.sampled_from(["numpy", "torch"]))
    n_conv = draw(st.integers(1, 2))
    convs = [(draw(st.integers(1, 6)), draw(st.sampled_from([1, 2, 3]) if flavour == "numpy" else st.just(3))) for _ in range(n_conv)]
    lo = 4 * n_conv + 4
    shape = (draw(st.integers(lo, 21)), draw(st.integers(lo, 21)), draw(st.integers(1, 3)))
    hidden = draw(st.lists(st.integers(1, 9), min_size=0, max_size=2))
    B = draw(st.integers(1, 5))
    mode = draw(st.sampled_from(["logit", "softmax_ce"]))
    return flavour, shape, convs, hidden, B, mode, draw(st.integers(0, 2 ** 16))


@settings(max_examples=25, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))   # fixed example set: no flaky runs
@given(nets())
@example(("torch", (12, 13, 1), [(4, 3), (2, 3)], [3], 1, "softmax_ce", 6710))    # a 2-channel cam with a tiny value range (found by a random run)
def test_shape_generic_parity(net):
    flavour, shape, convs, hidden, B, mode, seed = net
    mk = ocnn.NetConfig.numpy_flavour if flavour == "numpy" else ocnn.NetConfig.torch_flavour
    cfg = mk(shape, 2, convs, hidden, 0.01)
    p = ocnn.init_params(cfg, seed=seed, bias_std=0.05)
    x = ocnn.synth_images(B, shape, seed=seed + 1)
    eng = engine_from(cfg, p, max_batch=4)                      # B = 5 also exercises chunking
    cls, probs, logits, heat = eng.predict_explain(x, None, mode)
    o_cls, cache, A, dA, o_heat = oracle_heatmaps(cfg, p, x, None, mode, A_for_ties=None)
    lg = cache.logits.numpy()
    assert np.abs(_np(logits) - lg).max() <= 1e-4 * max(1.0, np.abs(lg).max())
    score = cache.probs.numpy() if cfg.head == "softmax" else lg
    margin = np.abs(score[:, 0] - score[:, 1])
    safe = margin > 1e-4
    assert np.array_equal(_np(cls)[safe], o_cls[safe])
    if cfg.pool_ties == "first" and safe.all():                  # (the tie-duplicating rule is precision-dependent: test_gpu_parity.py)
        # min-max normalisation divides by the cam's range: with 1-6 random channels that range can be tiny next to the
        # activations, so fp32 rounding is amplified -- bound the error relative to that conditioning (the fixed-shape tests gate
        # realistic networks at 1e-4)
        A = cache.conv_out[-1].numpy()
        cond = max(1.0, float(np.abs(A).max()) / max(1e-12, float(np.ptp(o_heat_raw(cfg, p, cache, o_cls if mode == "logit" else o_cls, mode)))))
        assert np.abs(_np(heat) - o_heat).max() <= 1e-4 * min(cond, 50.0) + 1e-4
    # batch-of-1 == slice of the batch
    c1, p1, l1, h1 = eng.predict_explain(x[B - 1:B], None, mode)
    assert np.array_equal(_np(l1)[0], _np(logits)[B - 1]) and np.array_equal(_np(h1)[0], _np(heat)[B - 1])
    eng.close()
