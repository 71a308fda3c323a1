"""GPU parity tests: the CUDA path (through the C-ABI) against the golden fixtures and the oracle.

Tolerances (north_star): classes exact; logits / probabilities / normalised heatmaps max-abs <= 1e-4 on
the fp32 path, <= 1e-2 on the bf16 tensor path.
"""
import numpy as np
import pytest
import torch

from util import engine_from, load_numpy_golden, load_torch_golden, oracle_heatmaps, ocnn

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name", ["ref_numpy_small", "ref_numpy_odd", "ref_numpy_ties", "ref_numpy_k5"])
def test_numpy_flavour_golden(name):
    import bcad_b200
    from bcad_b200 import _lib
    g, cfg, p, conv_idx, dense_idx = load_numpy_golden(name)
    eng = engine_from(cfg, p, max_batch=4, keep_all_activations=True)
    x = g["x"][None].astype(np.float32)
    cls, probs, logits = eng.predict(x)
    assert int(cls[0]) == int(g["pred_class"])
    np.testing.assert_allclose(_np(probs)[0], g["probs"], rtol=0, atol=FP32_TOL)
    for bi, li in enumerate(conv_idx):
        co = _np(eng.get_tensor(_lib.T_CONV_OUT, bi, 1))[0].reshape(g[f"conv_out{li}"].shape)
        po = _np(eng.get_tensor(_lib.T_POOL_OUT, bi, 1))[0].reshape(g[f"pool_out{li + 1}"].shape)
        np.testing.assert_allclose(co, g[f"conv_out{li}"], rtol=0, atol=FP32_TOL)
        np.testing.assert_allclose(po, g[f"pool_out{li + 1}"], rtol=0, atol=FP32_TOL)
    for j, li in enumerate(dense_idx):
        z = _np(eng.get_tensor(_lib.T_DENSE_Z, j, 1))[0]
        np.testing.assert_allclose(z, g[f"z{li}"], rtol=0, atol=FP32_TOL)
    for c in (0, 1):
        outs, d_in = eng.explain_backward(1, c, "softmax_ce", want_conv=range(len(conv_idx)), want_input=True)
        np.testing.assert_allclose(_np(d_in)[0], g[f"d_input_c{c}"], rtol=0, atol=FP32_TOL)
        for bi, li in enumerate(conv_idx):
            np.testing.assert_allclose(_np(outs[bi])[0], g[f"conv_act_grads{li}_c{c}"], rtol=0, atol=FP32_TOL)
    eng.close()


@pytest.mark.parametrize("name", ["ref_torch_small", "ref_torch_odd"])
def test_torch_flavour_golden(name):
    from bcad_b200 import _lib
    g, cfg, p = load_torch_golden(name)
    B = g["x"].shape[0]
    eng = engine_from(cfg, p, max_batch=4, keep_all_activations=True)
    cls, probs, logits = eng.predict(g["x"])
    assert np.array_equal(_np(cls), g["pred_class"])
    np.testing.assert_allclose(_np(logits), g["logits"], rtol=0, atol=FP32_TOL)
    np.testing.assert_allclose(_np(probs), g["probs"], rtol=0, atol=FP32_TOL)
    last = len(cfg.conv_layers) - 1
    A = _np(eng.get_tensor(_lib.T_CONV_OUT, last, B)).reshape(B, *cfg.shapes()[0][last][0])
    np.testing.assert_allclose(A.transpose(0, 3, 1, 2), g["A_last"], rtol=0, atol=FP32_TOL)
    for c in (0, 1):
        outs, _ = eng.explain_backward(B, c, "logit", want_conv=range(last + 1))
        for i in range(last + 1):
            np.testing.assert_allclose(_np(outs[i]).transpose(0, 3, 1, 2), g[f"dA{i}_logit_c{c}"], rtol=0, atol=FP32_TOL)
    eng.close()


CASES = [
    # (flavour, input_shape, conv_layers, hidden, batch, kind)
    ("numpy", (40, 36, 1), [(8, 3), (16, 3)], [32, 16], 5, "gauss"),
    ("torch", (40, 36, 1), [(8, 3), (16, 3)], [32, 16], 5, "gauss"),
    ("numpy", (33, 47, 3), [(5, 3), (7, 3)], [9], 3, "gauss"),
    ("torch", (32, 32, 4), [(32, 3), (64, 3)], [24, 8], 6, "gauss"),
    ("numpy", (48, 48, 1), [(8, 3), (16, 3)], [16], 4, "mammo"),      # exact-zero background => pool ties
    ("torch", (48, 48, 1), [(8, 3), (16, 3)], [16], 4, "mammo"),
    ("numpy", (30, 30, 2), [(6, 5), (8, 3), (10, 3)], [12], 3, "gauss"),   # three blocks, k=5
]


@pytest.mark.parametrize("flavour,shape,convs,hidden,B,kind", CASES)
@pytest.mark.parametrize("grad_mode", ["logit", "softmax_ce"])
def test_predict_explain_vs_oracle(flavour, shape, convs, hidden, B, kind, grad_mode):
    mk = ocnn.NetConfig.numpy_flavour if flavour == "numpy" else ocnn.NetConfig.torch_flavour
    cfg = mk(shape, 2, convs, hidden, 0.01)
    p = ocnn.init_params(cfg, seed=7, bias_std=0.05)
    x = ocnn.synth_images(B, shape, seed=123, kind=kind)
    eng = engine_from(cfg, p, max_batch=8)
    eng_chunked = engine_from(cfg, p, max_batch=4)               # B > max_batch for some cases => chunking
    from bcad_b200 import _lib
    from util import _switches_from, near_tie_windows
    last = len(convs) - 1
    for class_idx in (None, np.arange(B) % 2):
        cls, probs, logits, heat = eng.predict_explain(x, class_idx, grad_mode)
        # The pool-tie rule compares activations for exact equality: take the equality structure from the
        # device's own fp32 activations (last chunk is cached) and everything else from the float64 oracle.
        A_ties = _np(eng.get_tensor(_lib.T_CONV_OUT, last, B)).reshape(B, *cfg.shapes()[0][last][0])
        o_cls, cache, A, dA, o_heat = oracle_heatmaps(cfg, p, x, class_idx, grad_mode)
        sw_dev, sw_or = _switches_from(A_ties, cfg.pool_ties), cache.switches[-1].numpy()
        if not np.array_equal(sw_dev, sw_or):
            # differences are only allowed where float64 itself sees a near-tie (gap < 1e-6 relative)
            h2, w2 = A.shape[1] // 2, A.shape[2] // 2
            diff_win = (sw_dev != sw_or)[:, :2 * h2, :2 * w2].reshape(B, h2, 2, w2, 2, -1).any(axis=(2, 4))
            assert np.all(near_tie_windows(cache.conv_out[last].numpy())[diff_win]), "tie structure differs beyond fp32 rounding"
            o_cls, cache, A, dA, o_heat = oracle_heatmaps(cfg, p, x, class_idx, grad_mode, A_for_ties=A_ties)
        assert np.array_equal(_np(cls), o_cls)
        np.testing.assert_allclose(_np(logits), cache.logits.numpy(), rtol=0, atol=FP32_TOL)
        np.testing.assert_allclose(_np(probs), cache.probs.numpy(), rtol=0, atol=FP32_TOL)
        np.testing.assert_allclose(_np(heat), o_heat, rtol=0, atol=FP32_TOL)
        c2, p2, l2, h2 = eng_chunked.predict_explain(x, class_idx, grad_mode)      # chunked == unchunked
        assert np.array_equal(_np(c2), _np(cls))
        np.testing.assert_allclose(_np(l2), _np(logits), rtol=0, atol=1e-5)
        np.testing.assert_allclose(_np(h2), _np(heat), rtol=0, atol=1e-5)
    eng_chunked.close()
    # host-buffer C-ABI call == device-buffer call
    h_cls, h_probs, h_logits, h_heat = eng.predict_explain_host(x, None, grad_mode)
    cls, probs, logits, heat = eng.predict_explain(x, None, grad_mode)
    assert np.array_equal(h_cls, _np(cls))
    assert np.array_equal(h_heat, _np(heat)) and np.array_equal(h_logits, _np(logits))
    eng.close()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,K,h,w,H,W", [(3, 64, 31, 29, 64, 60), (2, 16, 62, 62, 125, 125), (4, 7, 9, 11, 20, 23),
                                           (2, 64, 128, 128, 256, 256), (1, 64, 125, 125, 256, 256)])
def test_standalone_tail_vs_oracle(dtype, B, K, h, w, H, W):
    import bcad_b200
    from oracle import gradcam as ogc
    rng = np.random.default_rng(5)
    A = torch.tensor(rng.standard_normal((B, h, w, K)), dtype=torch.float32).to(dtype)
    dA = torch.tensor(rng.standard_normal((B, h, w, K)) * 1e-3, dtype=torch.float32).to(dtype)
    want = ogc.gradcam_tail_nhwc(A.float().numpy(), dA.float().numpy(), (H, W))
    got = bcad_b200.gradcam_tail(A.cuda(), dA.cuda(), (H, W))
    np.testing.assert_allclose(_np(got), want, rtol=0, atol=FP32_TOL)


def test_tail_degenerate_maps():
    import bcad_b200
    A = torch.ones((2, 8, 8, 4), device="cuda")
    out = bcad_b200.gradcam_tail(A, A.clone(), (16, 16))              # constant cam: min == max => zeros
    assert float(out.abs().max()) == 0.0
    out = bcad_b200.gradcam_tail(A, -A, (16, 16))                     # all-negative cam => ReLU => zeros
    assert float(out.abs().max()) == 0.0


def test_overlay_vs_oracle():
    import bcad_b200
    from oracle import gradcam as ogc
    rng = np.random.default_rng(9)
    cam = rng.random((3, 40, 50), dtype=np.float32)
    cam[0, 0, 0], cam[0, 0, 1] = 0.0, 1.0
    img = rng.random((3, 40, 50), dtype=np.float32)
    ov, hu = bcad_b200.overlay(torch.from_numpy(img).cuda(), torch.from_numpy(cam).cuda())
    assert np.array_equal(_np(hu), ogc.heatmap_u8(cam))
    for i in range(3):
        want = ogc.show_cam_on_image(np.stack([img[i]] * 3, -1), cam[i], use_rgb=True).astype(np.int32)
        diff = np.abs(_np(ov)[i].astype(np.int32) - want)
        assert diff.max() <= 1 and (diff > 0).mean() < 0.01        # u8 truncation boundaries only


def test_full_size_canonical_net():
    """BASELINE configs at full size (256x256x1, conv 32/64, dense 256/128), both flavours: oracle parity on a
    few images + size-independent properties (batch slicing, chunking, host == device, range)."""
    for flavour in ("torch", "numpy"):
        mk = ocnn.NetConfig.numpy_flavour if flavour == "numpy" else ocnn.NetConfig.torch_flavour
        cfg = mk((256, 256, 1), 2, [(32, 3), (64, 3)], [256, 128], 0.01)
        p = ocnn.init_params(cfg, seed=7, bias_std=0.0)
        x = ocnn.synth_images(12, (256, 256, 1), seed=20251018)
        eng = engine_from(cfg, p, max_batch=8)
        cls, probs, logits, heat = eng.predict_explain(x, None, "logit")          # 12 > 8 => two chunks
        o_cls, cache, A, dA, o_heat = oracle_heatmaps(cfg, p, x[:3], None, "logit")
        assert np.array_equal(_np(cls)[:3], o_cls)
        np.testing.assert_allclose(_np(logits)[:3], cache.logits.numpy(), rtol=0, atol=FP32_TOL)
        np.testing.assert_allclose(_np(heat)[:3], o_heat, rtol=0, atol=FP32_TOL)
        h = _np(heat)
        assert h.min() >= 0.0 and h.max() <= 1.0 and np.all(h.reshape(12, -1).max(axis=1) > 0.999)
        c1, p1, l1, h1 = eng.predict_explain(x[5:6], None, "logit")               # batch-of-1 == slice of batch-of-N
        assert np.array_equal(_np(h1)[0], h[5]) and np.array_equal(_np(l1)[0], _np(logits)[5])
        hc, hp, hl, hh = eng.predict_explain_host(x, None, "logit")
        assert np.array_equal(hh, h) and np.array_equal(hc, _np(cls))
        eng.close()


def test_edge_inputs_and_errors():
    cfg = ocnn.NetConfig.torch_flavour((16, 16, 1), 2, [(4, 3), (8, 3)], [8])
    p = ocnn.init_params(cfg, seed=3, bias_std=0.1)
    eng = engine_from(cfg, p, max_batch=2)
    x = np.zeros((1, 16, 16, 1), np.float32)                                      # constant image: ties everywhere
    o_cls, cache, A, dA, o_heat = oracle_heatmaps(cfg, p, x, None, "logit")
    cls, probs, logits, heat = eng.predict_explain(x)
    assert int(cls[0]) == int(o_cls[0])
    np.testing.assert_allclose(_np(heat), o_heat, rtol=0, atol=FP32_TOL)
    with pytest.raises(ValueError):
        eng.predict(np.zeros((1, 15, 16, 1), np.float32))
    with pytest.raises(ValueError):
        eng.predict_explain(x, class_idx=[5])
    with pytest.raises(RuntimeError):
        eng.explain_backward(3, 0)                                                # no cached forward of that B
    eng.close()


def test_batchnorm_fold_identity_and_affine():
    cfg = ocnn.NetConfig.torch_flavour((16, 16, 1), 2, [(4, 3), (8, 3)], [8])
    p = ocnn.init_params(cfg, seed=3, bias_std=0.1)
    x = ocnn.synth_images(2, (16, 16, 1), seed=4)
    import bcad_b200
    from util import spec_from_cfg
    rng = np.random.default_rng(0)
    g, b, mu, var = rng.uniform(0.5, 1.5, 4), rng.normal(0, 0.1, 4), rng.normal(0, 0.1, 4), rng.uniform(0.5, 1.5, 4)
    eng = bcad_b200.Engine(spec_from_cfg(cfg), max_batch=2)
    eng.set_weights(p.conv_w, p.conv_b, p.dense_w, p.dense_b, batchnorm={0: (g, b, mu, var, 1e-5)})
    _, _, logits = eng.predict(x)
    sc = g / np.sqrt(var + 1e-5)
    p2 = ocnn.Params([p.conv_w[0] * sc[:, None, None, None], p.conv_w[1]], [(p.conv_b[0] - mu) * sc + b, p.conv_b[1]],
                     p.dense_w, p.dense_b)
    np.testing.assert_allclose(_np(logits), ocnn.forward(cfg, p2, x).logits.numpy(), rtol=0, atol=FP32_TOL)
    eng.close()


def test_sharded_engine_equals_single_gpu():
    """Batch sharding over the GPUs of the box: N-GPU result == 1-GPU result on the concatenated batch."""
    import bcad_b200
    from util import spec_from_cfg
    ngpu = torch.cuda.device_count()
    devices = list(range(min(ngpu, 4))) if ngpu > 1 else [0, 0]       # one GPU: two handles on it still exercise the split
    cfg = ocnn.NetConfig.torch_flavour((32, 32, 1), 2, [(8, 3), (16, 3)], [16], 0.01)
    p = ocnn.init_params(cfg, seed=2, bias_std=0.05)
    x = ocnn.synth_images(11, (32, 32, 1), seed=8)
    sh = bcad_b200.ShardedEngine(spec_from_cfg(cfg), devices, max_batch=4)
    sh.set_weights(p.conv_w, p.conv_b, p.dense_w, p.dense_b)
    cls, probs, logits, heat = sh.predict_explain_host(x, None, "logit")
    one = engine_from(cfg, p, max_batch=16)
    c1, p1, l1, h1 = one.predict_explain_host(x, None, "logit")
    assert np.array_equal(cls, c1)
    np.testing.assert_allclose(logits, l1, rtol=0, atol=1e-5)
    np.testing.assert_allclose(heat, h1, rtol=0, atol=1e-5)
    sh.close()
    one.close()


@pytest.mark.parametrize("flavour,kind", [("torch", "gauss"), ("numpy", "mammo")])
def test_gradcam_on_the_pre_activation_conv_output(flavour, kind):
    """target="conv_preact": what pytorch_grad_cam captures when it hooks ``model.convs[-1]`` of the reference's torch CNN -- the
    nn.Conv2d output BEFORE the functional F.leaky_relu (ADCNNM.py:76).  Oracle: z = a > 0 ? a : a / slope recovered from the
    post-activation map, dz = dA * LeakyReLU'(z), same tail.  fp32 engine, both pool-tie rules."""
    import bcad_b200
    from oracle import gradcam as ogc
    mk = ocnn.NetConfig.numpy_flavour if flavour == "numpy" else ocnn.NetConfig.torch_flavour
    cfg = mk((40, 36, 1), 2, [(8, 3), (16, 3)], [24, 12], 0.05)
    p = ocnn.init_params(cfg, seed=12, bias_std=0.05)
    x = ocnn.synth_images(5, (40, 36, 1), seed=9, kind=kind)
    eng = engine_from(cfg, p, max_batch=8)
    eng.set_explain_target("conv_preact")
    ci = np.array([0, 1, 1, 0, 1])
    cls, probs, logits, heat = eng.predict_explain(x, ci, "logit")
    cache = ocnn.forward(cfg, p, x)
    cag, _, _ = ocnn.backward(cfg, p, cache, ocnn.top_gradient(cache, ci, "logit"), through_input=False)
    A = cache.conv_out[1].numpy()
    slope = cfg.alpha_conv
    Z = np.where(A > 0, A, A / slope).astype(np.float32)
    dZ = (cag[1].numpy() * np.where(A > 0, 1.0, slope)).astype(np.float32)
    want = ogc.gradcam_tail_nhwc(Z, dZ, (40, 36))
    np.testing.assert_allclose(_np(heat), want, rtol=0, atol=FP32_TOL)
    eng.set_explain_target("conv")
    _, _, _, heat_act = eng.predict_explain(x, ci, "logit")
    _, _, _, _, o_heat = oracle_heatmaps(cfg, p, x, ci, "logit")
    np.testing.assert_allclose(_np(heat_act), o_heat, rtol=0, atol=FP32_TOL)
    assert np.abs(_np(heat_act) - want).max() > 1e-3                     # the two targets really differ
    eng.close()
    from util import spec_from_cfg
    if flavour == "torch":
        t = bcad_b200.Engine(spec_from_cfg(ocnn.NetConfig.torch_flavour((32, 32, 1), 2, [(32, 3), (64, 3)], [32], 0.01)), precision="fp16")
        with pytest.raises(ValueError, match="fp32 path"):
            t.set_explain_target("conv_preact")
        t.close()
