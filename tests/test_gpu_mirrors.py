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
