"""Training step on the GPU vs the reference-generated fixture and the oracle."""
import os

import numpy as np
import pytest
import torch

from util import GOLDEN, engine_from, ocnn

pytestmark = pytest.mark.gpu


def _cmp(got, want, tol, what):
    err = np.abs(got - want).max()
    assert err <= tol * max(1.0, np.abs(want).max()), f"{what}: err {err}"


def _fixture():
    g = np.load(os.path.join(GOLDEN, "ref_numpy_train.npz"))
    cfg = ocnn.NetConfig.numpy_flavour((12, 12, 2), 2, [(3, 3), (4, 3)], [6, 5], 0.01)
    p = ocnn.Params([g["W0"], g["W2"]], [g["b0"], g["b2"]], [g["W4"], g["W5"], g["W6"]], [g["b4"], g["b5"], g["b6"]])
    return g, cfg, p


def test_mean_gradients_and_sgd_clip_step_match_reference_fixture():
    """Reference _compute_sample_grads (averaged) and _apply_grads vs bcad_train_backward / bcad_apply_update."""
    g, cfg, p = _fixture()
    eng = engine_from(cfg, p, max_batch=4, keep_all_activations=True)
    x = torch.from_numpy(g["X"].astype(np.float32)).cuda()
    eng.predict(x)
    grads, loss = eng.train_backward(x, g["labels"])
    _cmp(loss.cpu().numpy(), g["losses"], 1e-5, "loss")
    u = eng.unpack_grads(grads)
    _cmp(u["conv_w"][0], g["grad0_dF"], 1e-5, "dF0")
    _cmp(u["conv_b"][0], g["grad0_db_conv"], 1e-5, "db0")
    _cmp(u["conv_w"][1], g["grad2_dF"], 1e-5, "dF2")
    _cmp(u["conv_b"][1], g["grad2_db_conv"], 1e-5, "db2")
    for j, li in enumerate((4, 5, 6)):
        _cmp(u["dense_w"][j], g[f"grad{li}_dW"], 1e-5, f"dW{li}")
        _cmp(u["dense_b"][j], g[f"grad{li}_db"], 1e-5, f"db{li}")
    eng.apply_update(grads, "sgd_clip", lr=float(g["lr"]), max_norm=5.0)
    cw, cb, dw, db = eng.get_weights()
    _cmp(cw[0], g["newW0"], 1e-5, "W0"); _cmp(cw[1], g["newW2"], 1e-5, "W2")
    _cmp(cb[0], g["newb0"], 1e-5, "b0"); _cmp(cb[1], g["newb2"], 1e-5, "b2")
    for j, li in enumerate((4, 5, 6)):
        _cmp(dw[j], g[f"newW{li}"], 1e-5, f"W{li}")
        _cmp(db[j], g[f"newb{li}"], 1e-5, f"b{li}")
    # the updated weights are the ones the next forward uses (incl. the refreshed dgrad copy)
    newp = ocnn.Params([g["newW0"], g["newW2"]], [g["newb0"], g["newb2"]], [g["newW4"], g["newW5"], g["newW6"]],
                       [g["newb4"], g["newb5"], g["newb6"]])
    _, probs, _ = eng.predict(x)
    _cmp(probs.cpu().numpy(), ocnn.forward(cfg, newp, g["X"]).probs.numpy(), 1e-5, "probs after update")
    eng.close()


@pytest.mark.parametrize("flavour,shape,convs,hidden,B", [
    ("torch", (20, 24, 1), [(8, 3), (16, 3)], [12], 6),
    ("numpy", (17, 15, 3), [(4, 3), (6, 3)], [7, 5], 5),
    ("torch", (64, 64, 1), [(32, 3), (64, 3)], [32, 16], 8),
])
def test_gradients_vs_oracle_and_adam(flavour, shape, convs, hidden, B):
    from oracle import train as otr
    mk = ocnn.NetConfig.numpy_flavour if flavour == "numpy" else ocnn.NetConfig.torch_flavour
    cfg = mk(shape, 2, convs, hidden, 0.01)
    p = ocnn.init_params(cfg, seed=3, bias_std=0.05)
    x = ocnn.synth_images(B, shape, seed=9)
    labels = np.arange(B) % 2
    eng = engine_from(cfg, p, max_batch=8, keep_all_activations=True)
    xd = torch.from_numpy(x).cuda()
    want_seq = []
    cur = p
    for step in range(2):                                     # two Adam steps: exercises m/v state and bias correction
        eng.predict(xd)
        grads, loss = eng.train_backward(xd, labels)
        want, wloss = otr.mean_grads(cfg, cur, x, labels)
        u = eng.unpack_grads(grads)
        for k in ("conv_w", "conv_b", "dense_w", "dense_b"):
            for a, b in zip(u[k], want[k]):
                _cmp(a, b, 2e-4, f"step {step} {k}")
        _cmp(loss.cpu().numpy(), wloss, 1e-4, "loss")
        want_seq.append(u)                                    # Adam checked on the gradients it was actually given
        eng.apply_update(grads, "adam", lr=1e-2)
        cur = otr.adam_steps(p, want_seq, lr=1e-2)
        got = eng.get_weights()
        for a, b in zip(got[0] + got[1] + got[2] + got[3], cur.conv_w + cur.conv_b + cur.dense_w + cur.dense_b):
            _cmp(a, b, 3e-4, f"weights after adam step {step}")
    eng.close()


def test_data_parallel_equals_full_batch():
    """Two 'ranks' (two handles) on half batches, gradients averaged == one handle on the full batch (the
    oracle for the all-reduce is 'gradient of the full batch on one device')."""
    cfg = ocnn.NetConfig.torch_flavour((16, 16, 1), 2, [(4, 3), (8, 3)], [8], 0.01)
    p = ocnn.init_params(cfg, seed=1, bias_std=0.05)
    x = torch.from_numpy(ocnn.synth_images(8, (16, 16, 1), seed=2)).cuda()
    labels = np.array([0, 1, 1, 0, 1, 0, 0, 1])
    full = engine_from(cfg, p, max_batch=8, keep_all_activations=True)
    full.predict(x)
    g_full, _ = full.train_backward(x, labels)
    halves = []
    for r in range(2):
        e = engine_from(cfg, p, max_batch=4, keep_all_activations=True)
        xs = x[4 * r:4 * r + 4]
        e.predict(xs)
        g, _ = e.train_backward(xs, labels[4 * r:4 * r + 4])
        halves.append(g.clone())
        e.close()
    avg = (halves[0] + halves[1]) / 2
    assert float((avg - g_full).abs().max()) <= 1e-6 * max(1.0, float(g_full.abs().max()))
    full.close()
