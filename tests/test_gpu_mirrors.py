"""The reference-shaped Python surface (CNNModel / ADCNNM / explainability / GRADCAM / ExplainableAI mirrors) on the GPU,
against the reference-generated fixtures and the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from util import GOLDEN, load_numpy_golden, load_torch_golden, oracle_heatmaps, ocnn

pytestmark = pytest.mark.gpu


def test_numpy_mirror_forward_predict_and_layer_caches():
    """CNNModel mirror == Classes/CNNModel.py on the reference fixture: predict, probs, layers[*] caches, switches."""
    from bcad_b200.CNNModel import CNNModel
    g, cfg, p, conv_idx, dense_idx = load_numpy_golden("ref_numpy_ties")
    np.random.seed(0)
    m = CNNModel(tuple(cfg.input_shape), 2, conv_layers=[list(c) for c in cfg.conv_layers], hidden_units=list(cfg.hidden_units),
                 dropout_rate=0.3, leaky_alpha=cfg.alpha_conv)
    for i, layer in enumerate(m.layers):
        if layer["type"] == "conv":
            layer["filters"], layer["biases"] = g[f"W{i}"], g[f"b{i}"]
        elif layer["type"] in ("dense", "output"):
            layer["weights"], layer["biases"] = g[f"W{i}"], g[f"b{i}"]
    cls, probs = m.predict(g["x"])
    assert int(cls) == int(g["pred_class"]) and probs.dtype == np.float64 and probs.shape == (2,)
    np.testing.assert_allclose(probs, g["probs"], rtol=0, atol=1e-5)
    for li in conv_idx:
        np.testing.assert_allclose(m.layers[li]["output"], g[f"conv_out{li}"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(m.layers[li + 1]["output"], g[f"pool_out{li + 1}"], rtol=0, atol=1e-5)
        assert np.array_equal(m.layers[li + 1]["switches"], g[f"switches{li + 1}"])
    for li in dense_idx:
        np.testing.assert_allclose(m.layers[li]["z"], g[f"z{li}"], rtol=0, atol=1e-5)
    assert m.layers[dense_idx[0]]["input"].shape == (m.layers[dense_idx[0]]["weights"].shape[1],)
    # batched entry points agree with the single-sample path
    X = np.stack([g["x"], g["x"][::-1].copy(), g["x"] * 0.5])
    classes, pb = m.predict_batch(X)
    assert classes[0] == cls and np.allclose(pb[0], probs, atol=1e-5)
    c2, p2, heat = m.predict_explain_batch(X)
    assert heat.shape == (3,) + tuple(cfg.input_shape[:2]) and np.array_equal(c2, classes)


def test_explainability_mirror_matches_reference_fixture(tmp_path):
    from bcad_b200.CNNModel import CNNModel
    from bcad_b200 import explainability as E
    g, cfg, p, conv_idx, dense_idx = load_numpy_golden("ref_numpy_small")
    m = CNNModel(tuple(cfg.input_shape), 2, conv_layers=[list(c) for c in cfg.conv_layers], hidden_units=list(cfg.hidden_units),
                 leaky_alpha=cfg.alpha_conv)
    for i, layer in enumerate(m.layers):
        if layer["type"] == "conv":
            layer["filters"], layer["biases"] = g[f"W{i}"], g[f"b{i}"]
        elif layer["type"] in ("dense", "output"):
            layer["weights"], layer["biases"] = g[f"W{i}"], g[f"b{i}"]
    m.forward(g["x"], training=False)
    for c in (0, 1):
        y = np.zeros(2, np.float32)
        y[c] = 1
        grads, d_input, cag = E.compute_backprops_for_explainability(m, y)
        assert len(grads) == len(m.layers) and set(cag.keys()) == set(conv_idx)
        np.testing.assert_allclose(d_input, g[f"d_input_c{c}"], rtol=0, atol=1e-5)
        for li in conv_idx:
            np.testing.assert_allclose(cag[li], g[f"conv_act_grads{li}_c{c}"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(E.saliency_map(d_input), g[f"saliency_c{c}"], rtol=0, atol=1e-4)
        for i, gr in enumerate(grads):                               # the weight gradients of the same backward (explainability.py:25,33,63)
            if m.layers[i]["type"] == "pool":
                assert gr is None
                continue
            assert set(gr.keys()) == ({"dF", "db_conv"} if m.layers[i]["type"] == "conv" else {"dW", "db"})
            for k, v in gr.items():
                assert v.dtype == np.float64 and v.shape == g[f"grad{i}_{k}_c{c}"].shape
                np.testing.assert_allclose(v, g[f"grad{i}_{k}_c{c}"], rtol=0, atol=1e-5)
    out = E.generate_dual_class_overlays(m, g["x"][..., :1].repeat(3, axis=-1) if False else g["x"], [0, 1], str(tmp_path / "xai"))
    assert set(out.keys()) == {0, 1}
    for c in (0, 1):
        assert os.path.exists(tmp_path / "xai" / f"overlay_class_{c}.png") and os.path.exists(tmp_path / "xai" / f"heatmap_class_{c}.png")
        assert out[c][0].dtype == np.uint8 and out[c][0].shape[:2] == g["x"].shape[:2]


def test_torch_mirror_forward_and_load_trained_model(tmp_path):
    from bcad_b200 import ADCNNM as A
    g, cfg, p = load_torch_golden("ref_torch_small")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    pth = tmp_path / "best_model.pth"
    torch.save(sd, pth)
    summary = {"dataset": {"input_shape": [int(v) for v in g["input_shape"]], "num_classes": 2},
               "model": {"conv_layers": [[int(a) for a in r] for r in g["conv_layers"]], "hidden_units": [int(v) for v in g["hidden"]],
                         "dropout_rate": 0.1}, "training": {"device": "cpu"}}
    js = tmp_path / "training_summary_advanced.json"
    js.write_text(json.dumps(summary))
    model = A.load_trained_model(str(js), str(pth))
    assert not model.training
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        out = model(x)                                           # app.py:586 call shape
    assert out.shape == (3, 2) and out.device == x.device
    np.testing.assert_allclose(out.numpy(), g["logits"], rtol=0, atol=1e-5)
    _, preds = torch.max(out, 1)                                  # app.py:589
    assert np.array_equal(preds.numpy(), g["pred_class"])
    np.testing.assert_allclose(torch.softmax(out, dim=1).numpy(), g["probs"], rtol=0, atol=1e-5)
    classes, logits, heat = model.predict_explain_batch(x.cuda(), class_idx=[0, 0, 0])
    from oracle import gradcam as ogc
    want = ogc.gradcam_tail(g["A_last"], g["dA1_logit_c0"], (16, 16))
    np.testing.assert_allclose(heat.cpu().numpy(), want, rtol=0, atol=1e-4)


def test_gradcam_mirror_and_explainable_ai(tmp_path):
    """generate_dual_class_gradcam_overlays_pytorch call surface (GRADCAM.py:31-81): dict of (overlay RGB u8, heat u8), PNGs."""
    from bcad_b200 import ADCNNM as A, GRADCAM as G
    from bcad_b200.ExplainableAI import ExplainableAI
    from oracle import gradcam as ogc
    cfg = ocnn.NetConfig.torch_flavour((32, 32, 1), 2, [(8, 3), (16, 3)], [16], 0.01)
    p = ocnn.init_params(cfg, seed=4, bias_std=0.05)
    model = A.CNNModel((32, 32, 1), 2, conv_layers=[(8, 3), (16, 3)], hidden_units=[16]).eval()
    model.load_state_dict(ocnn.params_to_state_dict(cfg, p))
    G.model = model
    rng = np.random.default_rng(1)
    img = (rng.random((32, 32)) * 255).astype(np.float32)
    out = G.generate_dual_class_gradcam_overlays_pytorch(img, [0, 1], str(tmp_path / "explainability"))
    assert set(out.keys()) == {0, 1}
    x = G.default_preprocess((img / 255.0).astype(np.float32), (32, 32, 1))
    for c in (0, 1):
        ov, hu = out[c]
        assert ov.shape == (32, 32, 3) and ov.dtype == np.uint8 and hu.shape == (32, 32) and hu.dtype == np.uint8
        assert os.path.exists(tmp_path / "explainability" / f"gradcam_overlay_class_{c}.png")
        assert os.path.exists(tmp_path / "explainability" / f"gradcam_heatmap_class_{c}.png")
        _, _, _, _, o_heat = oracle_heatmaps(cfg, p, x[None], np.array([c]), "logit")
        want_u8 = ogc.heatmap_u8(o_heat[0]).astype(np.int32)
        assert np.abs(hu.astype(np.int32) - want_u8).max() <= 1                      # u8 truncation boundary
        want_ov = ogc.show_cam_on_image(np.stack([img / 255.0] * 3, -1).astype(np.float32), o_heat[0]).astype(np.int32)
        assert (np.abs(ov.astype(np.int32) - want_ov) > 1).mean() < 0.02             # JET LUT steps at u8 boundaries
    pred = G.generate_dual_class_gradcam_overlays_pytorch(img, None, str(tmp_path / "explainability"), write_png=False)
    assert len(pred) == 1
    xai = ExplainableAI()
    hm = xai.generate_heatmap(model, x, 1)
    assert hm.shape == (32, 32) and xai.heatmap is hm and xai.last_conv_layer == 1 and xai.colormap == "jet"
    ov = xai.overlay_heatmap(img, hm)
    assert ov.shape == (32, 32, 3) and np.array_equal(ov, xai.visualize_prediction(img, hm))
    G.model = None
    with pytest.raises(RuntimeError):
        G.generate_dual_class_gradcam_overlays_pytorch(img)


def _numpy_mirror_from(g, cfg, **kw):
    from bcad_b200.CNNModel import CNNModel
    m = CNNModel(tuple(cfg.input_shape), 2, conv_layers=[list(c) for c in cfg.conv_layers], hidden_units=list(cfg.hidden_units),
                 leaky_alpha=cfg.alpha_conv, **kw)
    for i, layer in enumerate(m.layers):
        if layer["type"] == "conv":
            layer["filters"], layer["biases"] = g[f"W{i}"], g[f"b{i}"]
        elif layer["type"] in ("dense", "output"):
            layer["weights"], layer["biases"] = g[f"W{i}"], g[f"b{i}"]
    return m


def test_numpy_mirror_default_forward_is_the_reference_training_forward():
    """``model.forward(x)`` with the reference's defaults (training=True, dropout_rate=0.3: Classes/CNNModel.py:68,162,186-188):
    under the same seeded np.random stream the mirror draws the reference's dropout multipliers, so probs, the dense caches and
    the explainability backward of that forward equal the reference's own (fixture made by running the reference)."""
    from bcad_b200 import explainability as E
    g, cfg, p, conv_idx, dense_idx = load_numpy_golden("ref_numpy_forward_train")
    m = _numpy_mirror_from(g, cfg, dropout_rate=float(g["dropout_rate"]))
    np.random.seed(int(g["forward_seed"]))
    probs = m.forward(g["x"])                                        # training=True by default, as in the reference
    np.testing.assert_allclose(probs, g["probs"], rtol=0, atol=1e-5)
    for li in dense_idx:
        np.testing.assert_allclose(m.layers[li]["z"], g[f"z{li}"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(m.layers[li]["input"], g[f"dense_in{li}"], rtol=0, atol=1e-5)      # dropped activations
    eval_probs = m.forward(g["x"], training=False)
    assert np.abs(eval_probs - g["probs"]).max() > 1e-3              # (dropout did change the result)
    np.random.seed(int(g["forward_seed"]))
    m.forward(g["x"])
    grads, d_input, cag = E.compute_backprops_for_explainability(m, np.array([0.0, 1.0]))
    np.testing.assert_allclose(d_input, g["d_input_c1"], rtol=0, atol=1e-5)
    for i, gr in enumerate(grads):
        if gr is not None:
            for k, v in gr.items():
                np.testing.assert_allclose(v, g[f"grad{i}_{k}_c1"], rtol=0, atol=1e-5)


def test_numpy_mirror_caches_belong_to_their_forward():
    """The reference's layer caches belong to the forward() that filled them.  The mirror reads them lazily from a shared handle:
    if anything else ran on that handle in between, the forward is repeated (never another image's activations), and a layer
    dict kept from an earlier forward() refuses to answer."""
    from bcad_b200 import explainability as E
    g, cfg, p, conv_idx, dense_idx = load_numpy_golden("ref_numpy_small")
    m = _numpy_mirror_from(g, cfg)
    m.forward(g["x"], training=False)
    other = np.random.default_rng(0).standard_normal(g["x"].shape)
    m.engine.predict(other[None].astype(np.float32))                 # someone else uses the handle
    m.predict_batch(np.stack([other, other]))                        # ... and the batched entry point
    np.testing.assert_allclose(m.layers[conv_idx[0]]["output"], g[f"conv_out{conv_idx[0]}"], rtol=0, atol=1e-5)
    m.engine.predict(other[None].astype(np.float32))
    y = np.array([1.0, 0.0])
    _, d_input, _ = E.compute_backprops_for_explainability(m, y)
    np.testing.assert_allclose(d_input, g["d_input_c0"], rtol=0, atol=1e-5)
    # a getter captured during an earlier forward() must not answer with a later forward's data
    m.forward(g["x"], training=False)
    stale = dict.__getitem__(m.layers[dense_idx[0]], "z")
    m.forward(other, training=False)
    with pytest.raises(RuntimeError, match="earlier forward"):
        stale.fn()


def test_torch_mirror_train_mode_forward_applies_dropout():
    """ADCNNM.CNNModel is in train mode after construction (nn.Module default) and its forward then applies nn.Dropout
    (ADCNNM.py:62,72-78).  With injected multipliers the mirror equals the oracle's training forward; with drawn ones the
    result differs from eval mode and zeroes about p of the hidden activations' influence."""
    from bcad_b200 import ADCNNM as A
    cfg = ocnn.NetConfig.torch_flavour((32, 32, 1), 2, [(8, 3), (16, 3)], [16, 8], 0.01)
    p = ocnn.init_params(cfg, seed=4, bias_std=0.05)
    model = A.CNNModel((32, 32, 1), 2, conv_layers=[(8, 3), (16, 3)], hidden_units=[16, 8], dropout_rate=0.4, max_batch=4)
    model.load_state_dict(ocnn.params_to_state_dict(cfg, p))
    assert model.training
    x = torch.from_numpy(ocnn.synth_images(6, (32, 32, 1), seed=5))   # 6 > max_batch: two chunks, masks per chunk
    rng = np.random.default_rng(2)
    masks = ((rng.random((6, 24)) >= 0.4) / 0.6).astype(np.float32)
    got = model._forward_train(x, [0.4, 0.4], masks=masks)
    want = ocnn.forward(cfg, p, x.numpy(), dropout=[masks[:, :16], masks[:, 16:]]).logits.numpy()
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=1e-4)
    torch.manual_seed(0)
    out_train = model(x)
    assert out_train.shape == (6, 2) and out_train.device == x.device
    model.eval()
    out_eval = model(x)
    np.testing.assert_allclose(out_eval.numpy(), ocnn.forward(cfg, p, x.numpy()).logits.numpy(), rtol=0, atol=1e-4)
    assert np.abs(out_train.numpy() - out_eval.numpy()).max() > 1e-3


def test_host_call_from_two_threads_on_one_handle():
    """bcad_predict_explain_host is serialised per handle (shared streams / staging): two threads hammering one model -- the
    reference app's threaded requests on one module-level model (app.py:649-657) -- each get exactly their own results."""
    import threading
    from util import engine_from
    cfg = ocnn.NetConfig.torch_flavour((64, 48, 1), 2, [(32, 3), (64, 3)], [32, 16], 0.01)
    p = ocnn.init_params(cfg, seed=3, bias_std=0.05)
    eng = engine_from(cfg, p, precision="fp16", max_batch=32)
    xs = [ocnn.synth_images(n, (64, 48, 1), seed=s) for n, s in ((70, 1), (45, 2))]
    want = [eng.predict_explain_host(x, None, "logit") for x in xs]
    errs = []

    def work(i):
        try:
            for _ in range(6):
                c, pr, l, h = eng.predict_explain_host(xs[i], None, "logit")
                assert np.array_equal(c, want[i][0]) and np.array_equal(l, want[i][2]) and np.array_equal(h, want[i][3])
        except Exception as e:                                       # surfaced after join
            errs.append(e)

    ths = [threading.Thread(target=work, args=(i,)) for i in (0, 1)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    with pytest.raises(ValueError, match="class_idx"):
        eng.predict_explain_host(xs[0], np.full(70, 2, np.int32), "logit")
    from bcad_b200 import _lib
    import ctypes as C
    bad = np.full(70, -1, np.int32)
    out = np.empty((70, 64, 48), np.float32)
    rc = eng.lib.bcad_predict_explain_host(eng._h, C.c_void_p(xs[0].ctypes.data), 70, C.c_void_p(bad.ctypes.data), 0, None, None, None,
                                           C.c_void_p(out.ctypes.data))
    assert rc == _lib.ERR_INVALID and b"class_idx" in eng.lib.bcad_last_error()
    eng.close()


def test_gradcam_mirror_explains_an_image_of_any_size(tmp_path):
    """The reference hands GRADCAM.py a 512x512 image whatever the CNN was fed (app.py:649-657); pytorch_grad_cam then scales the
    low-resolution cam to THAT size (GRADCAM.py:46-64).  Mirror: the CNN sees cv2.resize(img / 255, model size), the heat-map is
    the cam resized straight to the image's size (one bilinear step, then the second min-max)."""
    import cv2
    from bcad_b200 import ADCNNM as A, GRADCAM as G
    from oracle import gradcam as ogc
    cfg = ocnn.NetConfig.torch_flavour((32, 32, 1), 2, [(8, 3), (16, 3)], [16], 0.01)
    p = ocnn.init_params(cfg, seed=4, bias_std=0.05)
    model = A.CNNModel((32, 32, 1), 2, conv_layers=[(8, 3), (16, 3)], hidden_units=[16]).eval()
    model.load_state_dict(ocnn.params_to_state_dict(cfg, p))
    rng = np.random.default_rng(7)
    img = (rng.random((80, 72)) * 255).astype(np.float32)            # not the model's size, not square
    out = G.generate_dual_class_gradcam_overlays_pytorch(img, [0, 1], str(tmp_path / "x"), model=model)
    img01 = (img / 255.0).astype(np.float32)
    x = G.default_preprocess(cv2.resize(img01, (32, 32), interpolation=cv2.INTER_LINEAR), (32, 32, 1))
    for c in (0, 1):
        ov, hu = out[c]
        assert ov.shape == (80, 72, 3) and hu.shape == (80, 72)
        cache = ocnn.forward(cfg, p, x[None])
        cag, _, _ = ocnn.backward(cfg, p, cache, ocnn.top_gradient(cache, np.array([c]), "logit"), through_input=False)
        want = ogc.gradcam_tail_nhwc(cache.conv_out[1].numpy().astype(np.float32), cag[1].numpy().astype(np.float32), (80, 72))[0]
        assert np.abs(hu.astype(np.int32) - ogc.heatmap_u8(want).astype(np.int32)).max() <= 1
        want_ov = ogc.show_cam_on_image(np.stack([img01] * 3, -1), want).astype(np.int32)
        assert (np.abs(ov.astype(np.int32) - want_ov) > 1).mean() < 0.02
    with pytest.raises(ValueError, match="grayscale"):
        G.generate_dual_class_gradcam_overlays_pytorch(np.zeros((8, 8, 3), np.float32), [0], str(tmp_path / "x"), model=model)


@pytest.mark.parametrize("shape,convs,hidden,B", [((32, 32, 1), [(8, 3), (16, 3)], [16], 5),
                                                   ((64, 48, 1), [(32, 3), (64, 3)], [32, 16], 70)])     # tensor path, > one host chunk
def test_gradcam_batch_call_uint8_in_uint8_out(shape, convs, hidden, B):
    """bcad_gradcam_overlays_host: a batch of 8-bit grey images in, show_cam_on_image overlays + heatmap_uint8 out, one call --
    equal to the single-image GRADCAM.py surface image by image (to the uint8 rounding boundary), and to the oracle."""
    from bcad_b200 import ADCNNM as A, GRADCAM as G
    from oracle import gradcam as ogc
    cfg = ocnn.NetConfig.torch_flavour(shape, 2, convs, hidden, 0.01)
    p = ocnn.init_params(cfg, seed=6, bias_std=0.05)
    model = A.CNNModel(shape, 2, conv_layers=convs, hidden_units=hidden, max_batch=32).eval()
    model.load_state_dict(ocnn.params_to_state_dict(cfg, p))
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, size=(B,) + shape[:2], dtype=np.uint8)
    if (shape[0] * shape[1]) & (shape[0] * shape[1] - 1) == 0:         # constant image: std = 0 -> x = 0 (the 1e-8 guard); only where numpy's
        imgs[1] = 17                                                    # float32 mean of N equal values is exact (N a power of two)
    cls, probs, ov, hu = G.generate_gradcam_overlays_batch(imgs, None, model=model)
    assert ov.shape == (B,) + shape[:2] + (3,) and ov.dtype == np.uint8 and hu.shape == (B,) + shape[:2] and hu.dtype == np.uint8
    x = np.stack([G.default_preprocess((im / 255.0).astype(np.float32), shape) for im in imgs])
    o_cls, cache, A_, dA, o_heat = oracle_heatmaps(cfg, p, x, None, "logit")
    assert np.array_equal(cls, o_cls)
    tol = 1 if not model.engine.uses_tensor_path or model.engine.precision == "fp16x3" else 3
    for i in range(B):
        assert np.abs(hu[i].astype(np.int32) - ogc.heatmap_u8(o_heat[i]).astype(np.int32)).max() <= tol, i
        want_ov = ogc.show_cam_on_image(np.stack([imgs[i] / 255.0] * 3, -1).astype(np.float32), o_heat[i]).astype(np.int32)
        assert (np.abs(ov[i].astype(np.int32) - want_ov) > 1).mean() < 0.02, i
    one = G.generate_dual_class_gradcam_overlays_pytorch(imgs[0].astype(np.float32), [int(cls[0])], "unused", model=model, write_png=False)
    assert np.abs(one[int(cls[0])][1].astype(np.int32) - hu[0].astype(np.int32)).max() <= 1
    c2, p2, ov2, hu2 = G.generate_gradcam_overlays_batch(imgs, 1, model=model, standardise=False)      # explicit class, img/255 as input
    x2 = (imgs / 255.0).astype(np.float32)[..., None]
    _, _, _, _, o_heat2 = oracle_heatmaps(cfg, p, x2, np.ones(B, np.int64), "logit")
    for i in (0, B - 1):
        assert np.abs(hu2[i].astype(np.int32) - ogc.heatmap_u8(o_heat2[i]).astype(np.int32)).max() <= tol
