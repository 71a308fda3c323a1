"""CPU-side tests: the C-ABI library loads and exports every symbol include/bcad.h declares, and the
host-side logic (specs, sharding, reference-shaped mirrors, persistence) behaves -- no GPU compute."""
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "bcad.h")).read()
    return sorted(set(re.findall(r"BCAD_API\s+[\w\s\*]+?\b(bcad_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge._build_module().build(force=False)
    import bcad_b200
    lib = bcad_b200._lib.load()
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libbcad.so does not export {name}"
    assert sorted(bcad_b200._lib.exported_symbols()) == declared, "ctypes table and header disagree"
    assert b"sm_100a" in lib.bcad_version()


def test_config_struct_layout_matches_header():
    import bcad_b200
    text = open(os.path.join(ROOT, "include", "bcad.h")).read()
    body = text[text.index("typedef struct bcad_config {"):text.index("} bcad_config;")]
    names = re.findall(r"\b(?:int32_t|float)\s+([^;]+);", body)
    flat = []
    for n in names:
        for part in n.split(","):
            flat.append(part.strip().split("[")[0])
    assert flat == [f[0] for f in bcad_b200._lib.Config._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_engine_fails_loudly_without_gpu():
    import bcad_b200
    spec = bcad_b200.NetSpec.torch_flavour((16, 16, 1), 2, [(4, 3)], [8])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bcad_b200.Engine(spec)


def test_netspec_shapes_match_reference_summary():
    import bcad_b200
    # Classes/CNNModel.py valid conv, floor pooling: 256 -> 254 -> 127 -> 125 -> 62
    s = bcad_b200.NetSpec.numpy_flavour((256, 256, 1), 2, [(32, 3), (64, 3)], [256, 128])
    shapes, flat = s.shapes()
    assert shapes == [((254, 254, 32), (127, 127, 32)), ((125, 125, 64), (62, 62, 64))] and flat == 246016
    t = bcad_b200.NetSpec.torch_flavour((256, 256, 1), 2, [(32, 3), (64, 3)], [256, 128])
    shapes, flat = t.shapes()
    assert shapes == [((256, 256, 32), (128, 128, 32)), ((128, 128, 64), (64, 64, 64))] and flat == 262144
    # literal trained shape (64,256,256) read as (H,W,C) (ADCNNM.py:42): conv1 is Conv2d(256->32)
    u = bcad_b200.NetSpec.torch_flavour((64, 256, 256), 2, [(32, 3), (64, 3)], [256, 128])
    assert u.shapes()[1] == 16 * 64 * 64


def test_shard_bounds_cover_batch_exactly():
    import bcad_b200
    for n in (0, 1, 7, 512, 8192, 8191):
        for world in (1, 2, 3, 4, 8):
            b = bcad_b200.shard_bounds(n, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_numpy_mirror_layers_structure(tmp_path):
    """CNNModel mirror: same layers list-of-dicts as Classes/CNNModel.py:88-157 and .npz round trip (:30-60,:530-555)."""
    from bcad_b200 import CNNModel as M
    np.random.seed(0)
    m = M.CNNModel((12, 12, 2), 2, conv_layers=[(3, 3), (4, 3)], hidden_units=[6, 5])
    types = [l["type"] for l in m.layers]
    assert types == ["conv", "pool", "conv", "pool", "dense", "dense", "output"]
    assert m.layers[0]["filters"].shape == (3, 3, 3, 2) and m.layers[0]["output_shape"] == (10, 10, 3)
    assert m.layers[1]["output_shape"] == (5, 5, 3) and m.layers[3]["output_shape"] == (1, 1, 4)
    assert m.layers[4]["weights"].shape == (6, 4) and m.layers[6]["weights"].shape == (2, 5)
    for k in ("input", "output"):
        assert m.layers[0][k] is None
    assert m.layers[1]["switches"] is None and m.layers[4]["z"] is None
    path = str(tmp_path / "trained_model" / "cnn.npz")
    m.save_model(path)
    data = np.load(path, allow_pickle=True)
    assert sorted(k for k in data.files if k != "config") == ["W0", "W2", "W4", "W5", "W6", "b0", "b2", "b4", "b5", "b6"]
    m2 = M.load_weights(M.CNNModel, path)
    for a, b in zip(m.layers, m2.layers):
        for k in ("filters", "weights", "biases"):
            if k in a:
                assert np.array_equal(a[k], b[k])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.forward(np.zeros((12, 12, 2)))                  # every forward runs on the device; there is nothing to fall back to


def test_golden_npz_loads_through_mirror_load_weights(tmp_path):
    """A checkpoint written in the reference's format (Classes/CNNModel.py:530-555) loads into the mirror."""
    from bcad_b200 import CNNModel as M
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_numpy_small.npz"))
    cfg = {"input_shape": [int(v) for v in g["input_shape"]], "num_classes": 2,
           "conv_layers": [[int(a) for a in r] for r in g["conv_layers"]], "hidden_units": [int(v) for v in g["hidden"]],
           "dropout_rate": 0.3}
    path = str(tmp_path / "ref.npz")
    np.savez(path, config=json.dumps(cfg), **{k: g[k] for k in g.files if re.fullmatch(r"[Wb]\d+", k)})
    m = M.load_weights(M.CNNModel, path)
    assert m.leaky_alpha == 0.01 and np.array_equal(m.layers[2]["filters"], g["W2"])


def test_torch_mirror_state_dict_keys():
    """ADCNNM mirror keeps the reference's parameter names (ADCNNM.py:43-70) so its checkpoints load."""
    from bcad_b200 import ADCNNM as A
    m = A.CNNModel((16, 16, 1), 2, conv_layers=[(4, 3), (8, 3)], hidden_units=[12, 6])
    keys = list(m.state_dict().keys())
    assert keys == ["convs.0.weight", "convs.0.bias", "convs.1.weight", "convs.1.bias",
                    "fc.0.weight", "fc.0.bias", "fc.3.weight", "fc.3.bias", "fc.6.weight", "fc.6.bias"]
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_torch_small.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    m.load_state_dict(sd)                                        # a reference state_dict loads unchanged
    assert m.fc[0].weight.shape == (12, 8 * 4 * 4)


def test_load_trained_model_errors(tmp_path):
    from bcad_b200 import ADCNNM as A
    cfg = {"dataset": {"input_shape": [16, 16, 1], "num_classes": 2},
           "model": {"conv_layers": [[4, 3]], "hidden_units": [8], "dropout_rate": 0.1}, "training": {"device": "cpu"}}
    jp = tmp_path / "summary.json"
    jp.write_text(json.dumps(cfg))
    with pytest.raises(FileNotFoundError):
        A.load_trained_model(str(jp), str(tmp_path / "missing.pth"))
    bad = tmp_path / "bad.pth"
    torch.save({"nope": torch.zeros(1)}, bad)
    with pytest.raises(RuntimeError):
        A.load_trained_model(str(jp), str(bad))


def test_explainability_rejects_non_onehot():
    from bcad_b200 import explainability as E
    with pytest.raises(ValueError):
        E._class_of([0.5, 0.5], 2)
    assert E._class_of([0, 1], 2) == 1


def test_u8_pixel_normalisation_is_one_rounding_whichever_way_the_reference_divides():
    """bcad_predict_explain_host_u8in computes x = float32(u8) / 255.0f with IEEE division.  The reference's callers divide in
    float32 (app.py:71 `torch.tensor(resized, dtype=torch.float32) / 255.0`) or in float64 followed by a float32 cast
    (GRADCAM.py:46 `img / 255.0` -> the float32 input tensor): for all 256 pixel values both give the same float32."""
    import torch
    px = np.arange(256, dtype=np.uint8)
    f32 = px.astype(np.float32) / np.float32(255.0)
    assert np.array_equal(f32, (px / 255.0).astype(np.float32))
    assert np.array_equal(f32, (torch.tensor(px, dtype=torch.float32) / 255.0).numpy())
    assert f32[0] == 0.0 and f32[255] == 1.0
