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


def test_dropout_masks_numpy_reference_semantics():
    """Reference forward(training=True) with dropout 0.4 + _compute_sample_grads (whose backward ignores the mask)."""
    g = np.load(os.path.join(GOLDEN, "ref_numpy_train_dropout.npz"))
    cfg = ocnn.NetConfig.numpy_flavour((12, 12, 2), 2, [(3, 3), (4, 3)], [8, 6], 0.01)
    p = ocnn.Params([g["W0"], g["W2"]], [g["b0"], g["b2"]], [g["W4"], g["W5"], g["W6"]], [g["b4"], g["b5"], g["b6"]])
    eng = engine_from(cfg, p, max_batch=4, keep_all_activations=True)
    x = torch.from_numpy(g["X"].astype(np.float32)).cuda()
    eng.set_dropout_masks(g["dropout_masks"], mask_backward=False)
    eng.predict(x)
    grads, loss = eng.train_backward(x, g["labels"])
    _cmp(loss.cpu().numpy(), g["losses"], 1e-5, "loss")
    u = eng.unpack_grads(grads)
    _cmp(u["conv_w"][0], g["grad0_dF"], 1e-5, "dF0")
    _cmp(u["conv_w"][1], g["grad2_dF"], 1e-5, "dF2")
    for j, li in enumerate((4, 5, 6)):
        _cmp(u["dense_w"][j], g[f"grad{li}_dW"], 1e-5, f"dW{li}")
        _cmp(u["dense_b"][j], g[f"grad{li}_db"], 1e-5, f"db{li}")
    # a forward of another batch size must refuse the stale masks; clearing them restores inference
    with pytest.raises(RuntimeError):
        eng.predict(x[:2])
    eng.set_dropout_masks(None)
    _, probs, _ = eng.predict(x)
    _cmp(probs.cpu().numpy(), ocnn.forward(cfg, p, g["X"]).probs.numpy(), 1e-5, "inference after clearing")
    eng.close()


def test_dropout_masks_autograd_semantics():
    from oracle import train as otr
    cfg = ocnn.NetConfig.torch_flavour((16, 20, 1), 2, [(4, 3), (8, 3)], [12, 10], 0.01)
    p = ocnn.init_params(cfg, seed=5, bias_std=0.05)
    x = ocnn.synth_images(6, (16, 20, 1), seed=6)
    labels = np.array([0, 1, 0, 1, 1, 0])
    rng = np.random.default_rng(0)
    mk = (rng.random((6, 22)) >= 0.3).astype(np.float32) / 0.7
    eng = engine_from(cfg, p, max_batch=6, keep_all_activations=True)
    eng.set_dropout_masks(mk, mask_backward=True)
    xd = torch.from_numpy(x).cuda()
    eng.predict(xd)
    grads, loss = eng.train_backward(xd, labels)
    want, wloss = otr.mean_grads(cfg, p, x, labels, dropout=[mk[:, :12], mk[:, 12:]], dropout_in_backward=True)
    u = eng.unpack_grads(grads)
    for k in ("conv_w", "conv_b", "dense_w", "dense_b"):
        for a, b in zip(u[k], want[k]):
            _cmp(a, b, 1e-4, k)
    _cmp(loss.cpu().numpy(), wloss, 1e-5, "loss")
    eng.close()


def _toy_dataset(n, shape, seed):
    """Two classes told apart by which half of the image is bright."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n,) + shape).astype(np.float32) * 0.3
    y = rng.integers(0, 2, n)
    for i in range(n):
        if y[i]:
            X[i, : shape[0] // 2] += 1.5
        else:
            X[i, shape[0] // 2:] += 1.5
    return X, y


def test_numpy_mirror_train_batch_matches_oracle_and_train_learns(capsys):
    from bcad_b200.CNNModel import CNNModel
    from oracle import train as otr
    shape = (16, 16, 1)
    X, y = _toy_dataset(24, shape, 1)
    Y = np.eye(2)[y]
    np.random.seed(7)
    m = CNNModel(shape, 2, conv_layers=[(4, 3), (6, 3)], hidden_units=[10, 8], dropout_rate=0.25, max_batch=8)
    cfg = ocnn.NetConfig.numpy_flavour(shape, 2, [(4, 3), (6, 3)], [10, 8], 0.01)
    p0 = ocnn.Params([l["filters"].copy() for l in m._conv_layers()], [l["biases"].copy() for l in m._conv_layers()],
                     [l["weights"].copy() for l in m._dense_layers()], [l["biases"].copy() for l in m._dense_layers()])
    # one batch: the mirror draws the dropout multipliers from np.random in the reference's order
    np.random.seed(99)
    mk = np.stack([np.concatenate([(np.random.rand(u) > 0.25).astype(np.float32) / 0.75 for u in (10, 8)]) for _ in range(8)])
    np.random.seed(99)
    loss_sum = m.train_batch(X[:8], Y[:8], lr=0.05)
    want, wloss = otr.mean_grads(cfg, p0, X[:8], y[:8], dropout=[mk[:, :10], mk[:, 10:]], dropout_in_backward=False)
    p1 = otr.sgd_clip_step(p0, want, 0.05)
    assert abs(loss_sum - wloss.sum()) < 1e-4 * max(1.0, wloss.sum())
    m._pull_weights()
    for l, w in zip(m._conv_layers(), p1.conv_w):
        _cmp(l["filters"], w, 1e-5, "filters after one batch")
    for l, w in zip(m._dense_layers(), p1.dense_w):
        _cmp(l["weights"], w, 1e-5, "weights after one batch")
    # the whole loop: learns the toy problem, restores the best weights, keeps layers[] in sync with the device
    m.train(X, Y, X, Y, epochs=12, lr=0.05, batch_size=8)
    out = capsys.readouterr().out
    assert "[TRAIN] Best accuracy" in out
    acc = m.get_training_metrics(X, Y, verbose=False)
    assert acc >= 0.9 and abs(acc - max(m.epoch_accuracy)) < 1e-9
    cw, _, dw, _ = m.engine.get_weights()
    _cmp(cw[0], m._conv_layers()[0]["filters"], 1e-6, "layers[] == device")
    _cmp(dw[0], m._dense_layers()[0]["weights"], 1e-6, "layers[] == device")


def test_torch_mirror_train_model(tmp_path):
    from bcad_b200 import ADCNNM as A
    shape = (16, 16, 1)
    X, y = _toy_dataset(32, shape, 2)
    Xt, yt = torch.from_numpy(X), torch.from_numpy(y).long()
    loader = [(Xt[i:i + 8], yt[i:i + 8]) for i in range(0, 32, 8)]
    torch.manual_seed(0)
    m = A.CNNModel(shape, 2, conv_layers=[(4, 3), (8, 3)], hidden_units=[16, 8], dropout_rate=0.2)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    path = str(tmp_path / "best.pth")
    hist, best = A.train_model(m, loader, loader, epochs=15, lr=5e-3, save_path=path)
    assert len(hist) == 15 and hist[-1]["loss"] < hist[0]["loss"] and best >= 0.9
    after = m.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before)          # parameters came back from the device
    m.eval()
    lg = m(Xt)                                                                # uses the trained device weights
    assert float((lg.argmax(1).cpu() == yt).float().mean()) >= 0.9
    saved = torch.load(path)
    assert set(saved) == set(after)
    m2 = A.CNNModel(shape, 2, conv_layers=[(4, 3), (8, 3)], hidden_units=[16, 8], dropout_rate=0.2).eval()
    m2.load_state_dict(saved)
    assert float((m2(Xt).argmax(1).cpu() == yt).float().mean()) >= 0.9


def test_torch_mirror_train_model_on_tensor_cores(tmp_path):
    """train_model(..., tensor_cores=True): the mirror's training loop with the tcgen05 kernels (dropout on) learns the toy problem too,
    and refuses a network without an eligible block."""
    from bcad_b200 import ADCNNM as A
    shape = (32, 32, 1)
    X, y = _toy_dataset(32, shape, 2)
    Xt, yt = torch.from_numpy(X), torch.from_numpy(y).long()
    loader = [(Xt[i:i + 8], yt[i:i + 8]) for i in range(0, 32, 8)]
    torch.manual_seed(0)
    m = A.CNNModel(shape, 2, conv_layers=[(32, 3), (64, 3)], hidden_units=[128, 16], dropout_rate=0.2)
    hist, best = A.train_model(m, loader, loader, epochs=12, lr=2e-3, save_path=str(tmp_path / "best.pth"), tensor_cores=True)
    assert len(hist) == 12 and hist[-1]["loss"] < hist[0]["loss"] and best >= 0.9
    small = A.CNNModel((16, 16, 1), 2, conv_layers=[(4, 3), (8, 3)], hidden_units=[16, 8], dropout_rate=0.2)
    Xs = torch.zeros((8, 16, 16, 1))
    with pytest.raises(ValueError, match="no conv block"):
        A.train_model(small, [(Xs, yt[:8])], [(Xs, yt[:8])], epochs=1, tensor_cores=True)


def test_nccl_data_parallel_step_matches_full_batch():
    """2 ranks over NCCL (needs 2 GPUs; the driver's 1-GPU run skips it -- the gloo test covers the collective on CPU)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(root, "tools", "train_dp_check.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DP-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("shape,pad,B,hidden", [
    ((64, 64, 1), 1, 6, [32, 16]),        # second block on 32x32 maps
    ((61, 61, 1), 0, 5, [32, 16]),        # valid conv, odd maps: 29x29 -> 27x27 (a last row without a partner, "full" dgrad padding of 2)
    ((40, 200, 1), 1, 4, [32, 16]),       # 100-pixel rows: two 64-pixel halves in the weight gradient
    ((64, 64, 1), 1, 6, [128, 16]),       # + the first dense layer's weight / input gradient GEMMs on tcgen05 (128 units: one half)
    ((64, 64, 1), 1, 5, [256, 32]),       #   256 units (the canonical width): two halves, four streamed stages per tile
])
def test_fast_training_matches_the_fp32_kernels(shape, pad, B, hidden):
    """bcad_set_fast_training: the 32 -> 64 conv block's forward, input gradient and weight gradient on tcgen05 (split operands) against
    the reference-pinned fp32 kernels of the same handle: logits, loss and every gradient tensor to 1e-3 of the tensor's largest entry
    (measured ~1e-5), then two Adam steps stay together."""
    cfg = ocnn.NetConfig(shape, 2, [(32, 3), (64, 3)], hidden, 0.01, 0.01, pad, "chw", "first", "logits")
    p = ocnn.init_params(cfg, seed=3, bias_std=0.05)
    x = torch.from_numpy(ocnn.synth_images(B, shape, seed=9)).cuda()
    labels = np.arange(B) % 2
    ref = engine_from(cfg, p, max_batch=8, keep_all_activations=True)
    fast = engine_from(cfg, p, max_batch=8, keep_all_activations=True)
    fast.set_fast_training(True)
    for step in range(2):
        _, _, l_ref = ref.predict(x)
        _, _, l_fast = fast.predict(x)
        _cmp(l_fast.cpu().numpy(), l_ref.cpu().numpy(), 1e-3, f"step {step} logits")
        g_ref, loss_ref = ref.train_backward(x, labels)
        g_fast, loss_fast = fast.train_backward(x, labels)
        _cmp(loss_fast.cpu().numpy(), loss_ref.cpu().numpy(), 1e-3, "loss")
        u_ref, u_fast = ref.unpack_grads(g_ref), fast.unpack_grads(g_fast)
        for k in ("conv_w", "conv_b", "dense_w", "dense_b"):
            for i, (a, b) in enumerate(zip(u_fast[k], u_ref[k])):
                err = np.abs(a - b).max()
                assert err <= 1e-3 * max(1e-12, np.abs(b).max()), f"step {step} {k}[{i}]: err {err} vs max {np.abs(b).max()}"
        ref.apply_update(g_ref, "adam", lr=1e-3)
        fast.apply_update(g_fast, "adam", lr=1e-3)
    ref.close()
    fast.close()


def test_fast_training_rejects_networks_without_an_eligible_block():
    cfg = ocnn.NetConfig.torch_flavour((20, 24, 1), 2, [(8, 3), (16, 3)], [12], 0.01)
    p = ocnn.init_params(cfg, seed=3)
    eng = engine_from(cfg, p, max_batch=4, keep_all_activations=True)
    with pytest.raises(ValueError, match="no conv block"):
        eng.set_fast_training(True)
    eng.close()
