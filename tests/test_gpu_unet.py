"""Tiny U-Net encoder front on the GPU vs the reference-generated fixtures and the oracle."""
import os

import numpy as np
import pytest

from util import GOLDEN

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.mark.parametrize("name,c", [("ref_unet_small", 1), ("ref_unet_odd", 2)])
def test_unet_front_vs_reference_fixture(name, c):
    from bcad_b200 import unet as U
    from oracle import unet as ou
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    ks = ou.draw_kernels(c, int(g["seed"]))
    c1 = U.conv2d(g["x"], ks[0])
    assert c1.shape == g["c1"].shape
    assert np.all(c1[:, -2:] == 0) and np.all(c1[:, :, -2:] == 0)
    assert _rel(c1, g["c1"]) <= 1e-5
    assert _rel(U.max_pool(U.relu(g["c1"])), g["p1"]) <= 1e-6
    bn = U.tiny_unet(g["x"], ks)
    assert bn.shape == g["bn"].shape and _rel(bn, g["bn"]) <= 1e-4       # randn(std 1) kernels: values in the 100s
    assert _rel(U.average_pool(g["bn"], 3), g["avg3"]) <= 1e-6
    np.random.seed(int(g["seed"]))                                         # the reference's own entry point + RNG stream
    assert _rel(U.tiny_unet_numpy(g["x"]), g["bn"]) <= 1e-4


def test_unet_front_feeds_the_cnn_full_size():
    """BASELINE config 3 shape flow: 256x256x1 -> tiny U-Net (67x67x64) -> average_pool(3) (22x22x64) -> CNN + Grad-CAM."""
    import torch
    from bcad_b200 import unet as U
    from oracle import unet as ou
    from oracle import cnn as ocnn
    from util import engine_from, oracle_heatmaps
    x = ocnn.synth_images(3, (256, 256, 1), seed=5)
    ks = [k * s for k, s in zip(ou.draw_kernels(1, 7), (0.3, 0.08, 0.06))]        # scaled so activations stay O(1)
    bn = U.tiny_unet(x, ks)
    want = ou.tiny_unet(x.astype(np.float64), ks)
    assert bn.shape == (3, 67, 67, 64) and _rel(bn, want) <= 1e-4
    feat = U.average_pool(bn, 3)
    assert feat.shape == (3, 22, 22, 64)
    cfg = ocnn.NetConfig.torch_flavour((22, 22, 64), 2, [(32, 3), (64, 3)], [64, 32], 0.01)
    p = ocnn.init_params(cfg, seed=7, bias_std=0.05)
    eng = engine_from(cfg, p, max_batch=4)
    cls, probs, logits, heat = eng.predict_explain(feat.astype(np.float32), None, "logit")
    o_cls, cache, A, dA, o_heat = oracle_heatmaps(cfg, p, feat.astype(np.float32), None, "logit")
    assert np.array_equal(cls.cpu().numpy(), o_cls)
    assert np.abs(logits.cpu().numpy() - cache.logits.numpy()).max() <= 1e-4 * max(1.0, np.abs(cache.logits.numpy()).max())
    assert np.abs(heat.cpu().numpy() - o_heat).max() <= 1e-4
    eng.close()


def test_process_bottleneck_features_matches_reference():
    """app.py:466-489 outputs recorded from the reference (tests/golden/ref_bottleneck.npz): bit-exact."""
    import os
    import numpy as np
    import torch
    from util import GOLDEN
    from bcad_b200 import bottleneck as bn
    g = np.load(os.path.join(GOLDEN, "ref_bottleneck.npz"))
    got0 = bn.process_bottleneck_features(torch.from_numpy(g["feat0"]), resize_shape=(8, 8))        # tensor: [C,H,W]
    got1 = bn.process_bottleneck_features(g["feat1"], resize_shape=(11, 7))                         # ndarray CHW, odd sizes
    got2 = bn.process_bottleneck_features(g["feat2"], resize_shape=(9, 10))                         # ndarray already HWC
    for got, want in ((got0, g["out0"]), (got1, g["out1"]), (got2, g["out2"])):
        assert got.shape == want.shape and got.dtype == np.float32
        assert np.array_equal(got, want), float(np.abs(got - want).max())
    # <= 4 channels take OpenCV's double-coordinate path (the Grad-CAM map's): checked against the oracle
    from oracle import gradcam as ogc
    rng = np.random.default_rng(5)
    f = rng.standard_normal((3, 41, 29)).astype(np.float32)
    want = ogc.process_bottleneck_features(f, (13, 17))
    assert np.abs(bn.process_bottleneck_features(f, (13, 17)) - want).max() <= 1e-6
    # batched entry point, 64 channels, 256 -> 32 like the app
    fb = rng.standard_normal((2, 64, 256, 256)).astype(np.float32)
    out = bn.resize_batch(fb, (32, 32)).cpu().numpy()
    assert np.array_equal(out[1], ogc.process_bottleneck_features(fb[1], (32, 32)))


def test_unet_front_on_tensor_cores_vs_reference_fixture():
    """bcad_unet_* (conv2 / conv3 as tcgen05 implicit GEMMs, fp16 operands) against the reference's own tiny_unet_numpy output
    and average_pool (fixture made by running Classes/unet.py): 16-bit tolerance, quirk rows / columns exactly zero."""
    from bcad_b200 import unet as U
    from oracle import unet as ou
    g = np.load(os.path.join(GOLDEN, "ref_unet_small.npz"))
    ks = ou.draw_kernels(1, int(g["seed"]))
    front = U.UnetFront(16, 16, ks, max_batch=2)
    bn = front.forward(g["x"], avg_pool=0).cpu().numpy()
    assert bn.shape == g["bn"].shape == (2, 7, 7, 64) and front.out_shape(0) == (7, 7, 64)
    assert np.all(bn[:, -2:] == 0) and np.all(bn[:, :, -2:] == 0)
    assert _rel(bn, g["bn"]) <= 1e-2
    avg = front.forward(g["x"], avg_pool=3).cpu().numpy()
    assert avg.shape == g["avg3"].shape and _rel(avg, g["avg3"]) <= 1e-2
    assert front.launch_count == 10
    front.close()
    with pytest.raises(ValueError, match="multiples of 4"):
        U.UnetFront(18, 16)


@pytest.mark.parametrize("H,W,B,scales", [(256, 256, 5, (0.3, 0.08, 0.06)),      # BASELINE cfg 3 shape, activations O(1)
                                          (256, 256, 3, (1.0, 1.0, 1.0)),         # the reference's raw randn kernels: values in the 1000s
                                          (64, 48, 7, (0.3, 0.08, 0.06)),
                                          (8, 12, 3, (0.3, 0.1, 0.1)),            # smallest maps: 4x6 -> 3x4
                                          (40, 520, 2, (0.3, 0.08, 0.06))])       # 260-pixel rows: conv2 walks three 128-pixel segments
def test_unet_front_on_tensor_cores_vs_oracle(H, W, B, scales):
    import torch
    from bcad_b200 import unet as U
    from oracle import unet as ou
    from oracle import cnn as ocnn
    x = ocnn.synth_images(B, (H, W, 1), seed=5)
    ks = [k * s for k, s in zip(ou.draw_kernels(1, 7), scales)]
    front = U.UnetFront(H, W, ks, max_batch=4)                                   # B > max_batch: chunked
    want = ou.tiny_unet(x.astype(np.float64), ks)
    bn = front.forward(torch.from_numpy(x).cuda(), avg_pool=0).cpu().numpy()
    assert bn.shape == want.shape == (B, H // 4 + 3, W // 4 + 3, 64)
    assert np.isfinite(bn).all() and _rel(bn, want) <= 1e-2
    assert np.all(bn[:, -2:] == 0) and np.all(bn[:, :, -2:] == 0)
    # the border row / column of the pooled second map is the part computed outside the tensor-core kernel: check it on its own
    h3, w3 = H // 4 + 1, W // 4 + 1
    assert _rel(bn[:, h3 - 2:h3, :w3], want[:, h3 - 2:h3, :w3]) <= 1e-2 and _rel(bn[:, :h3, w3 - 2:w3], want[:, :h3, w3 - 2:w3]) <= 1e-2
    for pool in (3, 5):
        if min(h3 + 2, w3 + 2) >= pool:
            got = front.forward(x, avg_pool=pool).cpu().numpy()
            ref = ou.average_pool(want, pool)
            assert got.shape == ref.shape and _rel(got, ref) <= 1e-2
    front.close()
