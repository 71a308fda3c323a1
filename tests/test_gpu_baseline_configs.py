"""The BASELINE.json configurations that are parity cases rather than bench lines, at their full sizes, through
size-independent properties (+ the oracle on a few images).

cfg 1: 245 synthetic 256x256 images in batches of 32 (7 x 32 + 21: a ragged last batch)
cfg 3: U-Net front -> CNN -> Grad-CAM -> overlay   (the shape flow is in test_gpu_unet.py; here the overlay stage is added)
cfg 4: 8192 images, sharded / chunked
"""
import numpy as np
import pytest
import torch

from util import compare_all_images, engine_from, oracle_heatmaps, ocnn

pytestmark = pytest.mark.gpu


def _canonical():
    cfg = ocnn.NetConfig.torch_flavour((256, 256, 1), 2, [(32, 3), (64, 3)], [256, 128], 0.01)
    return cfg, ocnn.init_params(cfg, seed=7, bias_std=0.0)


@pytest.mark.parametrize("precision,tol", [("fp16", 1e-2), ("fp32", 1e-4)])
def test_cfg1_245_images_in_batches_of_32(precision, tol):
    cfg, p = _canonical()
    x = ocnn.synth_images(245, (256, 256, 1), seed=20251018)
    eng32 = engine_from(cfg, p, precision=precision, max_batch=32)          # 7 x 32 + 21
    cls, probs, logits, heat = eng32.predict_explain(x, None, "logit")
    cls, logits, heat = cls.cpu().numpy(), logits.cpu().numpy(), heat.cpu().numpy()
    assert heat.shape == (245, 256, 256) and np.isfinite(heat).all() and heat.min() >= 0.0 and heat.max() <= 1.0
    assert np.all(heat.reshape(245, -1).max(axis=1) > 0.999)               # every map is min-max normalised
    # batching must not matter: the ragged 21-image batch and a single image, run on their own, reproduce their slices bit for bit
    c21, p21, l21, h21 = eng32.predict_explain(x[224:], None, "logit")
    assert np.array_equal(cls[224:], c21.cpu().numpy()) and np.array_equal(logits[224:], l21.cpu().numpy())
    assert np.array_equal(heat[224:], h21.cpu().numpy())
    c1, p1, l1, h1 = eng32.predict_explain(x[100:101], None, "logit")
    assert np.array_equal(logits[100:101], l1.cpu().numpy()) and np.array_equal(heat[100:101], h1.cpu().numpy())
    # a handle with another max_batch splits fc1's K range differently (another fixed summation order), and in the 16-bit mode a
    # handle sized for 256 images runs the fused conv kernel while one sized for 32 runs the two conv kernels (the pooled
    # first-block map is rounded to fp16 at a different point): equal to the rounding of the mode
    eng_big = engine_from(cfg, p, precision=precision, max_batch=256)
    c2, p2, l2, h2 = eng_big.predict_explain(x, None, "logit")
    rnd = 2e-3 if precision == "fp16" else 2e-5
    assert np.abs(logits - l2.cpu().numpy()).max() <= rnd * max(1.0, np.abs(logits).max())
    assert np.array_equal(cls, c2.cpu().numpy())
    # every image of both handles against the oracle (heat-maps of two handles are compared through the oracle, not with each
    # other: a rounding-level change of a hidden pre-activation near 0 legitimately picks the other LeakyReLU branch)
    tau = 1.5e-2 if precision == "fp16" else 1e-4
    for eng in (eng32, eng_big):
        (r,) = compare_all_images(cfg, p, x, eng, [(None, "logit")], tau=tau)
        assert r["mask_violations"] == 0
        assert r["cls_equal"].all(), np.flatnonzero(~r["cls_equal"])
        assert r["logit_err"].max() <= tol * max(1.0, r["logit_absmax"].max())
        assert r["heat_err"].max() <= tol, np.sort(r["heat_err"])[-5:]
    eng32.close()
    eng_big.close()


def test_2048_images_every_one_against_the_oracle():
    """North-star gate on the benchmarked configuration (cfg 2 network, 256x256x1): 2048 distinct images, BOTH top-gradient
    modes, BOTH tensor-core modes (fp16 with small-margin refinement = the benchmarked path; fp16x3 = the fp32-grade default of
    the drop-in mirrors), every image compared with the float64 oracle -- classes identical on all of them, logits and
    heat-maps within the mode's tolerance on all of them.  No image is dropped, no flip allowance: the only latitude is the
    LeakyReLU' branch of hidden units whose oracle |z| is below the path's stated pre-activation error bound
    (util.compare_all_images), and the number of images that needed it is printed."""
    cfg, p = _canonical()
    n = 2048
    x = ocnn.synth_images(n, (256, 256, 1), seed=424242)
    e16 = engine_from(cfg, p, precision="fp16", max_batch=512)
    ex3 = engine_from(cfg, p, precision="fp16x3", max_batch=512)
    assert e16.refine_margin > 0
    bounds = {"fp16": (1e-2, 1e-2, 1.5e-2), "fp16x3": (2e-4, 5e-4, 5e-4)}       # logits, heat-maps, tau
    res = compare_all_images(cfg, p, x, [(e16, bounds["fp16"][2]), (ex3, bounds["fp16x3"][2])],
                             [(None, "logit"), (np.arange(n) % 2, "softmax_ce")], tau=None)
    refined, overflow = e16.refine_stats()
    for precision, per_mode in zip(("fp16", "fp16x3"), res):
        tol_l, tol_h, _ = bounds[precision]
        for name, r in zip(("predicted class / logit", "alternating class / softmax-CE"), per_mode):
            print(f"[{precision}] {name}: class mismatches {int((~r['cls_equal']).sum())}/{n}, max logit err {r['logit_err'].max():.3e}, "
                  f"max heat err {r['heat_err'].max():.3e}, images with a near-kink branch override {int(r['overridden'].sum())}, "
                  f"smallest margin {r['margin'].min():.3e}")
            assert r["mask_violations"] == 0
            assert np.array_equal(r["cls_equal"], np.ones(n, bool))
            assert r["logit_err"].max() <= tol_l * max(1.0, r["logit_absmax"].max())
            assert r["heat_err"].max() <= tol_h
    print(f"[fp16] images re-run at fp32 grade (top-2 logit gap < {e16.refine_margin}): {refined} of {2 * n} forwarded, overflowed {overflow}")
    assert overflow == 0 and refined > 0        # ~1 % of random-init images have a top-2 gap below the margin
    e16.close()
    ex3.close()


def test_cfg4_8192_images_chunked_host_call():
    """8192 images through the host-buffer call (128-image transfer chunks, max_batch 512): 128 copies of 64 distinct images --
    every copy must reproduce the first bit for bit, whatever chunk / position it lands in, and match a 64-image call."""
    cfg, p = _canonical()
    base = ocnn.synth_images(64, (256, 256, 1), seed=99)
    x = np.ascontiguousarray(np.broadcast_to(base[None], (128, 64, 256, 256, 1)).reshape(8192, 256, 256, 1))
    eng = engine_from(cfg, p, precision="fp16", max_batch=512)
    cls, probs, logits, heat = eng.predict_explain_host(x, None, "logit", heat_dtype=np.uint8)
    ref_c, ref_p, ref_l, ref_h = eng.predict_explain_host(base, None, "logit", heat_dtype=np.uint8)
    assert np.array_equal(cls.reshape(128, 64), np.broadcast_to(ref_c, (128, 64)))
    assert np.array_equal(logits.reshape(128, 64, 2), np.broadcast_to(ref_l, (128, 64, 2)))
    assert np.array_equal(heat.reshape(128, 64, 256, 256), np.broadcast_to(ref_h, (128, 64, 256, 256)))
    # and the device-buffer call over 512-image chunks agrees with the host call
    xd = torch.from_numpy(x[:1024]).cuda()
    c2, p2, l2, h2 = eng.predict_explain(xd, None, "logit")
    assert np.array_equal(c2.cpu().numpy(), cls[:1024]) and np.array_equal(l2.cpu().numpy(), logits[:1024])
    assert np.array_equal((h2 * 255).to(torch.uint8).cpu().numpy(), heat[:1024])
    eng.close()


def test_cfg3_overlay_stage_after_gradcam():
    """show_cam_on_image + heatmap_uint8 on the Grad-CAM maps of the canonical network vs the oracle (GRADCAM.py:67,70)."""
    import bcad_b200
    from oracle import gradcam as ogc
    cfg, p = _canonical()
    x = ocnn.synth_images(4, (256, 256, 1), seed=3)
    eng = engine_from(cfg, p, precision="fp32", max_batch=4)
    cls, probs, logits, heat = eng.predict_explain(x, None, "logit")
    img01 = torch.from_numpy((x[..., 0] - x.min()) / (x.max() - x.min())).cuda()            # grey image in [0,1]
    ov, hu = bcad_b200.overlay(img01, heat)
    heat_np, img_np = heat.cpu().numpy(), img01.cpu().numpy()
    for i in range(4):
        want_ov = ogc.show_cam_on_image(np.repeat(img_np[i][..., None], 3, axis=-1), heat_np[i])
        diff = np.abs(ov[i].cpu().numpy().astype(np.int16) - want_ov.astype(np.int16))
        assert diff.max() <= 1 and (diff > 0).mean() < 0.01                  # +-1 only at float32 rounding boundaries
        assert np.array_equal(hu[i].cpu().numpy(), ogc.heatmap_u8(heat_np[i]))
    eng.close()
