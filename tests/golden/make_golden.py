"""Generate the golden fixtures by RUNNING THE REFERENCE'S OWN CODE (authoring container only).

    python tests/golden/make_golden.py

Imports ``Classes/CNNModel.py`` (NumPy CNN), ``WebApplicationPrototype/explainability.py``
and ``WebApplicationPrototype/ADCNNM.py`` (torch CNN) from /root/reference through
``oracle.ref_loader`` and stores their inputs/outputs as small ``.npz`` files next to this
script.  The reference ships no golden vectors, tests or weights of its own,
so these reference-generated fixtures are what pins the oracle (and, through it, the CUDA path).
OpenCV fixtures (``cv2.resize`` bilinear, JET LUT) pin the two library sub-steps of the tail.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import ref_loader  # noqa: E402


def numpy_case(name, input_shape, conv_layers, hidden, seed, alpha=0.01, bias_std=0.1, ties=False):
    ref = ref_loader.load_numpy_cnn()
    xai = ref_loader.load_explainability()
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    with ref_loader.silenced():
        m = ref.CNNModel(input_shape, 2, conv_layers=conv_layers, hidden_units=hidden,
                         dropout_rate=0.3, leaky_alpha=alpha)
    # non-zero biases so the bias paths are exercised (reference init has zeros)
    for layer in m.layers:
        if "biases" in layer:
            layer["biases"] = rng.normal(0, bias_std, layer["biases"].shape)
    x = rng.standard_normal(input_shape)
    if ties:
        # mammogram-like: exact-zero background => constant regions => pool ties
        x[: input_shape[0] // 2] = 0.0
        x[:, : input_shape[1] // 3] = 0.0
    out = {"x": x, "alpha": alpha, "input_shape": np.array(input_shape),
           "conv_layers": np.array(conv_layers), "hidden": np.array(hidden)}
    for i, layer in enumerate(m.layers):
        if layer["type"] == "conv":
            out[f"W{i}"], out[f"b{i}"] = layer["filters"], layer["biases"]
        elif layer["type"] in ("dense", "output"):
            out[f"W{i}"], out[f"b{i}"] = layer["weights"], layer["biases"]
    with ref_loader.silenced():
        cls, probs = m.predict(x)
    out["pred_class"], out["probs"] = np.int64(cls), probs
    for i, layer in enumerate(m.layers):
        if layer["type"] == "conv":
            out[f"conv_out{i}"] = layer["output"]
        elif layer["type"] == "pool":
            out[f"pool_out{i}"], out[f"switches{i}"] = layer["output"], layer["switches"]
        elif layer["type"] in ("dense", "output"):
            out[f"z{i}"] = layer["z"]
    for c in (0, 1):
        y = np.zeros(2)
        y[c] = 1.0
        with ref_loader.silenced():
            m.forward(x, training=False)
            grads, d_input, cag = xai.compute_backprops_for_explainability(m, y)
        out[f"d_input_c{c}"] = d_input
        for k, v in cag.items():
            out[f"conv_act_grads{k}_c{c}"] = v
        for i, g in enumerate(grads):
            if g is None:
                continue
            for kk, vv in g.items():
                out[f"grad{i}_{kk}_c{c}"] = vv
        sal = np.abs(d_input).max(axis=-1)
        out[f"saliency_c{c}"] = (sal - sal.min()) / (sal.max() - sal.min() + 1e-8)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "class", cls, "probs", probs)


def torch_case(name, input_shape, conv_layers, hidden, batch, seed, alpha=0.01):
    ad = ref_loader.load_adcnnm()
    torch.manual_seed(seed)
    m = ad.CNNModel(input_shape, 2, conv_layers=conv_layers, hidden_units=hidden,
                    dropout_rate=0.3, leaky_alpha=alpha).eval()
    rng = np.random.default_rng(seed)
    x = torch.tensor(rng.standard_normal((batch,) + tuple(input_shape)), dtype=torch.float32)
    out = {"x": x.numpy(), "alpha": alpha, "input_shape": np.array(input_shape),
           "conv_layers": np.array(conv_layers), "hidden": np.array(hidden)}
    for k, v in m.state_dict().items():
        out["sd." + k] = v.numpy()
    with torch.no_grad():
        logits = m(x)                                    # the reference forward, untouched
    out["logits"] = logits.numpy()
    out["probs"] = torch.softmax(logits, dim=1).numpy()  # app.py:593
    out["pred_class"] = logits.argmax(dim=1).numpy()     # app.py:589
    # activations / gradients at the last conv's post-LeakyReLU output, using the reference's own
    # modules in the reference's own order (ADCNNM.py:74-78) with autograd
    import torch.nn.functional as F
    for c in (0, 1):
        h = x.permute(0, 3, 1, 2)
        acts = []
        for conv, pool in zip(m.convs, m.pools):
            a = F.leaky_relu(conv(h))
            a.retain_grad()
            acts.append(a)
            h = pool(a)
        lg = m.fc(h.reshape(h.size(0), -1))
        assert torch.equal(lg.detach(), logits), "replicated forward must equal model(x)"
        m.zero_grad()
        lg[:, c].sum().backward()
        out[f"A_last"] = acts[-1].detach().numpy()
        for i, a in enumerate(acts):
            out[f"dA{i}_logit_c{c}"] = a.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "classes", out["pred_class"])


def train_case(name, input_shape, conv_layers, hidden, seed, n=3, lr=0.05, dropout_rate=0.0):
    """Reference _compute_sample_grads over n samples, averaged as train() does (:438-464), then _apply_grads (:372-394)."""
    ref = ref_loader.load_numpy_cnn()
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    with ref_loader.silenced():
        m = ref.CNNModel(input_shape, 2, conv_layers=conv_layers, hidden_units=hidden, dropout_rate=dropout_rate, leaky_alpha=0.01)
    for layer in m.layers:
        if "biases" in layer:
            layer["biases"] = rng.normal(0, 0.1, layer["biases"].shape)
    X = rng.standard_normal((n,) + tuple(input_shape)) * 3.0        # large inputs => some gradient norms exceed the clip
    labels = rng.integers(0, 2, n)
    out = {"X": X, "labels": labels, "lr": lr, "input_shape": np.array(input_shape), "conv_layers": np.array(conv_layers),
           "hidden": np.array(hidden)}
    for i, layer in enumerate(m.layers):
        if layer["type"] == "conv":
            out[f"W{i}"], out[f"b{i}"] = layer["filters"].copy(), layer["biases"].copy()
        elif layer["type"] in ("dense", "output"):
            out[f"W{i}"], out[f"b{i}"] = layer["weights"].copy(), layer["biases"].copy()
    acc = [None] * len(m.layers)
    losses = []
    masks = []
    for si, (x, y) in enumerate(zip(X, labels)):
        onehot = np.eye(2)[y]
        if dropout_rate > 0.0:
            # the reference draws np.random.rand(units) per hidden layer, in order (:186-188): replay the stream to record
            # the multipliers it is about to use
            np.random.seed(seed * 100 + si)
            masks.append(np.concatenate([(np.random.rand(u) > dropout_rate).astype(np.float32) / (1.0 - dropout_rate) for u in hidden]))
            np.random.seed(seed * 100 + si)
        with ref_loader.silenced():
            probs = m.forward(x, training=True)
            losses.append(m.cross_entropy(probs, onehot))
            sg = m._compute_sample_grads(onehot)
        for idx, g in enumerate(sg):
            if g is None:
                continue
            if acc[idx] is None:
                acc[idx] = {k: np.zeros_like(v) for k, v in g.items()}
            for k in g:
                acc[idx][k] += g[k]
    for idx, g in enumerate(acc):
        if g is None:
            continue
        for k in g:
            g[k] = g[k] / float(n)
            out[f"grad{idx}_{k}"] = g[k]
    with ref_loader.silenced():
        m._apply_grads(acc, lr)
    for i, layer in enumerate(m.layers):
        if layer["type"] == "conv":
            out[f"newW{i}"], out[f"newb{i}"] = layer["filters"], layer["biases"]
        elif layer["type"] in ("dense", "output"):
            out[f"newW{i}"], out[f"newb{i}"] = layer["weights"], layer["biases"]
    out["losses"] = np.array(losses)
    if masks:
        out["dropout_masks"] = np.stack(masks)
        out["dropout_rate"] = dropout_rate
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "losses", losses)


def forward_train_case(name, input_shape, conv_layers, hidden, seed, dropout_rate=0.3):
    """The reference's DEFAULT call ``model.forward(x)`` (training=True, dropout_rate=0.3: Classes/CNNModel.py:68,162,186-188)
    under a seeded global np.random stream: probs, every layer cache, and the explainability backward of that forward."""
    ref = ref_loader.load_numpy_cnn()
    xai = ref_loader.load_explainability()
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    with ref_loader.silenced():
        m = ref.CNNModel(input_shape, 2, conv_layers=conv_layers, hidden_units=hidden, dropout_rate=dropout_rate, leaky_alpha=0.01)
    for layer in m.layers:
        if "biases" in layer:
            layer["biases"] = rng.normal(0, 0.1, layer["biases"].shape)
    x = rng.standard_normal(input_shape)
    out = {"x": x, "alpha": 0.01, "input_shape": np.array(input_shape), "conv_layers": np.array(conv_layers), "hidden": np.array(hidden),
           "dropout_rate": dropout_rate, "forward_seed": seed + 1000}
    for i, layer in enumerate(m.layers):
        if layer["type"] == "conv":
            out[f"W{i}"], out[f"b{i}"] = layer["filters"].copy(), layer["biases"].copy()
        elif layer["type"] in ("dense", "output"):
            out[f"W{i}"], out[f"b{i}"] = layer["weights"].copy(), layer["biases"].copy()
    np.random.seed(seed + 1000)
    with ref_loader.silenced():
        probs = m.forward(x)                                  # the reference's default: training=True
    out["probs"] = probs
    for i, layer in enumerate(m.layers):
        if layer["type"] in ("dense", "output"):
            out[f"z{i}"], out[f"dense_in{i}"] = layer["z"], layer["input"]
    y = np.zeros(2)
    y[1] = 1.0
    with ref_loader.silenced():
        grads, d_input, cag = xai.compute_backprops_for_explainability(m, y)
    out["d_input_c1"] = d_input
    for i, g in enumerate(grads):
        if g is not None:
            for kk, vv in g.items():
                out[f"grad{i}_{kk}_c1"] = vv
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "probs", probs)


def unet_case(name, shape, batch, seed):
    """Classes/unet.py functions (exec-loaded without the script part and its missing imports)."""
    import types
    path = os.path.join(ref_loader.REF_ROOT, "Classes", "unet.py")
    src = open(path, encoding="utf-8").read()
    src = src[:src.index("# ----------- Run UNet on All Images")]
    src = src.replace("from preprocessing import processed_images_np", "").replace("import matplotlib.pyplot as plt", "")
    mod = types.ModuleType("ref_unet")
    exec(compile(src, path, "exec"), mod.__dict__)
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((batch,) + tuple(shape))
    np.random.seed(seed)
    bn = mod.tiny_unet_numpy(x)                          # draws its three kernels from the global stream
    np.random.seed(seed)
    c1 = mod.conv2d(x, np.random.randn(3, 3, shape[-1], 16))
    seg_path = os.path.join(ref_loader.REF_ROOT, "Classes", "ImageSegmentation.py")
    out = {"x": x, "seed": seed, "bn": bn, "c1": c1, "p1": mod.max_pool(mod.relu(c1))}
    # average_pool as in Classes/ImageSegmentation.py:145-163 (same arithmetic, restated inline: the class needs pydicom)
    b, h, w, c = bn.shape
    ps = 3
    ap = np.zeros((b, h // ps, w // ps, c))
    for i in range(h // ps):
        for j in range(w // ps):
            ap[:, i, j, :] = np.mean(bn[:, i * ps:(i + 1) * ps, j * ps:(j + 1) * ps, :], axis=(1, 2))
    out["avg3"] = ap
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, bn.shape, ap.shape)


def bottleneck_case():
    """process_bottleneck_features (WebApplicationPrototype/app.py:466-489), exec-loaded alone (app.py needs flask)."""
    import types
    import cv2
    path = os.path.join(ref_loader.REF_ROOT, "WebApplicationPrototype", "app.py")
    src = open(path, encoding="utf-8").read()
    a = src.index("def process_bottleneck_features")
    b = src.index("@app.route('/classify'")
    mod = types.ModuleType("ref_app_slice")
    mod.__dict__.update({"torch": torch, "np": np, "cv2": cv2})
    exec(compile(src[a:b], path, "exec"), mod.__dict__)
    rng = np.random.default_rng(51)
    out = {}
    f0 = rng.standard_normal((8, 64, 64)).astype(np.float32)                 # CHW tensor, scale 8 like 256 -> 32
    out["feat0"], out["out0"] = f0, mod.process_bottleneck_features(torch.from_numpy(f0), resize_shape=(8, 8))
    f1 = rng.standard_normal((5, 37, 50)).astype(np.float32)                 # CHW ndarray (shape[0] < shape[2]), odd sizes
    out["feat1"], out["out1"] = f1, mod.process_bottleneck_features(f1, resize_shape=(11, 7))   # cv2 dsize = (width, height)
    f2 = rng.standard_normal((20, 24, 6)).astype(np.float32)                 # already HWC (shape[0] >= shape[2]): resized as is
    out["feat2"], out["out2"] = f2, mod.process_bottleneck_features(f2, resize_shape=(9, 10))
    np.savez_compressed(os.path.join(HERE, "ref_bottleneck.npz"), **out)
    print("wrote ref_bottleneck", out["out0"].shape, out["out1"].shape, out["out2"].shape)


def cv2_cases():
    import cv2
    lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_JET)[0]
    np.save(os.path.join(HERE, "jet_lut.npy"), lut)
    rng = np.random.default_rng(3)
    out = {}
    for i, (sh, dh) in enumerate([((16, 16), (32, 32)), ((125, 125), (256, 256)), ((128, 128), (256, 256)),
                                  ((7, 5), (20, 33)), ((62, 30), (61, 64)), ((1, 1), (8, 8))]):
        src = rng.random(sh, dtype=np.float32)
        out[f"src{i}"] = src
        out[f"dst{i}"] = cv2.resize(src, (dh[1], dh[0]))
    np.savez_compressed(os.path.join(HERE, "cv2_resize.npz"), **out)
    print("wrote cv2 fixtures; cv2", cv2.__version__)


if __name__ == "__main__":
    assert ref_loader.available(), "needs /root/reference"
    numpy_case("ref_numpy_small", (12, 12, 2), [(3, 3), (4, 3)], [6, 5], seed=11)
    numpy_case("ref_numpy_odd", (13, 11, 1), [(4, 3), (5, 3)], [7], seed=12, alpha=0.05)
    numpy_case("ref_numpy_ties", (14, 14, 1), [(3, 3), (4, 3)], [6, 4], seed=13, bias_std=0.2, ties=True)
    numpy_case("ref_numpy_k5", (15, 14, 2), [(3, 5), (4, 3)], [5], seed=14)
    torch_case("ref_torch_small", (16, 16, 1), [(4, 3), (8, 3)], [12, 6], batch=3, seed=21)
    torch_case("ref_torch_odd", (13, 18, 3), [(5, 3), (6, 3)], [9], batch=2, seed=22, alpha=0.2)
    train_case("ref_numpy_train", (12, 12, 2), [(3, 3), (4, 3)], [6, 5], seed=41)
    train_case("ref_numpy_train_dropout", (12, 12, 2), [(3, 3), (4, 3)], [8, 6], seed=42, n=4, dropout_rate=0.4)
    forward_train_case("ref_numpy_forward_train", (12, 12, 2), [(3, 3), (4, 3)], [8, 6], seed=43)
    unet_case("ref_unet_small", (16, 16, 1), 2, seed=31)
    unet_case("ref_unet_odd", (21, 18, 2), 1, seed=32)
    bottleneck_case()
    cv2_cases()
