"""CPU restatement of the Grad-CAM tail and overlay -- TEST INFRASTRUCTURE.

PARITY UNPINNED for the composition: the arithmetic is in the third-party package
``pytorch-grad-cam`` (PyPI ``grad-cam``; the reference pins no version and it is not
vendored under /root/reference).  What is restated here is that package's published
algorithm (``base_cam.py`` / ``grad_cam.py`` / ``utils/image.py``), anchored on the
reference's call sites:

* ``GRADCAM.py:53``  ``GradCAM(model=model, target_layers=[...])``
* ``GRADCAM.py:64``  ``cam(input_tensor, targets=[ClassifierOutputTarget(c)])[0]``
      w = dA.mean(axis=(2,3)); cam = (w[:,:,None,None]*A).sum(1); cam = max(cam,0)
      per image: cam -= min; cam /= (1e-7+max); cv2.resize(float32, (W,H))  [INTER_LINEAR]
      (one target layer: the max(.,0)/mean over layers is the identity)
      per image again: cam -= min; cam /= (1e-7+max)                -> float32 [B,H,W]
* ``GRADCAM.py:67``  ``show_cam_on_image(img_rgb, cam, use_rgb=True)``
      heat = applyColorMap(uint8(255*cam), JET) -> RGB -> /255 ; out = 0.5*heat + 0.5*img
      out /= out.max() ; uint8(255*out)
* ``GRADCAM.py:70``  ``(cam*255).astype(uint8)``  (truncation)
* ``explainability.py:71-78``  saliency: |d_input|.max(-1) -> min-max(1e-8) -> uint8 (truncate)

The two OpenCV sub-steps are pinned against OpenCV itself (tests/golden/cv2_resize.npz,
tests/golden/jet_lut.npy): ``bilinear_resize`` below is the pure-NumPy form that runs on the
GPU box without depending on cv2, and the tests check it equals ``cv2.resize``.
"""
from __future__ import annotations

import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def bilinear_resize(src: np.ndarray, out_h: int, out_w: int, coords_f32: bool = False) -> np.ndarray:
    """``cv2.resize(src.astype(float32), (out_w, out_h), interpolation=INTER_LINEAR)`` over the last two axes.

    coords_f32: OpenCV 4.13 has two float32 code paths (measured, tests/golden/ref_bottleneck.npz + cv2_resize.npz):
    images with <= 4 channels keep the source coordinate in double (default here: the Grad-CAM map is 1 channel); images with
    more channels (the 64-channel bottleneck features of app.py:466-489) use the classic loop
    ``fx = (float)((dx+0.5)*scale - 0.5); sx = cvFloor(fx); fx -= sx`` -- coordinate and weight in float32.

    Half-pixel centres ``s = (d+0.5)*(n_src/n_dst) - 0.5`` in double; ``i0=floor(s)``; weight =
    float32(s - i0); indices clamped to [0, n-1] with weight 0 (border replicate); separable; float32.
    Horizontal pass first then vertical, as OpenCV's resize does.
    """
    src = np.asarray(src, dtype=np.float32)
    h, w = src.shape[-2:]

    def coords(n_src, n_dst):
        # OpenCV 4.13 (wide-row path): the source coordinate is kept in double, the weight is the
        # fractional part rounded to float32 (measured against cv2: <= 1 ulp for every size tried;
        # float32 coordinates are 1.6e-6 off for non-power-of-two scales).
        s = (np.arange(n_dst, dtype=np.float64) + 0.5) * (np.float64(n_src) / np.float64(n_dst)) - 0.5
        if coords_f32:
            s = s.astype(np.float32)
        i0 = np.floor(s).astype(np.int64)
        f = (s - i0.astype(s.dtype)).astype(np.float32)
        # OpenCV: if i0 < 0 -> i0 = 0, f = 0 ; if i0 >= n-1 -> i0 = n-1, f = 0
        lo = i0 < 0
        hi = i0 >= n_src - 1
        f = np.where(lo | hi, np.float32(0), f)
        i0 = np.clip(i0, 0, n_src - 1)
        i1 = np.clip(i0 + 1, 0, n_src - 1)
        return i0, i1, f

    x0, x1, fx = coords(w, out_w)
    y0, y1, fy = coords(h, out_h)
    one = np.float32(1)
    rows = src[..., :, x0] * (one - fx) + src[..., :, x1] * fx               # [..., h, out_w]
    rows = rows.astype(np.float32)
    out = rows[..., y0, :] * (one - fy)[:, None] + rows[..., y1, :] * fy[:, None]
    return out.astype(np.float32)


def process_bottleneck_features(feat, resize_shape=(32, 32)) -> np.ndarray:
    """app.py:466-489: a [C,H,W] feature map (tensor, or ndarray with shape[0] < shape[2]) -> HWC, then
    ``cv2.resize(feat, resize_shape, INTER_LINEAR)`` (dsize = (width, height)) -> float32 [h, w, C]."""
    feat = np.asarray(feat, dtype=np.float32)
    chw = feat if feat.shape[0] < feat.shape[2] else feat.transpose(2, 0, 1)
    ow, oh = resize_shape
    return np.ascontiguousarray(bilinear_resize(chw, oh, ow, coords_f32=chw.shape[0] > 4).transpose(1, 2, 0))


def gradcam_tail(A: np.ndarray, dA: np.ndarray, out_hw, resize=bilinear_resize) -> np.ndarray:
    """A, dA: float32 [B,K,h,w] (NCHW, as the hooks of pytorch_grad_cam hold them).
    Returns float32 [B,H,W] in [0,1]."""
    A = np.asarray(A, dtype=np.float32)
    dA = np.asarray(dA, dtype=np.float32)
    H, W = out_hw
    weights = np.mean(dA, axis=(2, 3))                                        # alpha_k
    cam = (weights[:, :, None, None] * A).sum(axis=1)
    cam = np.maximum(cam, 0)
    out = []
    for img in cam:
        img = img - np.min(img)
        img = img / (1e-7 + np.max(img))
        img = resize(np.float32(img), H, W)
        out.append(img)
    out = np.float32(out)
    res = []
    for img in out:                                                           # second scale_cam_image
        img = img - np.min(img)
        img = img / (1e-7 + np.max(img))
        res.append(img)
    return np.float32(res)


def gradcam_tail_nhwc(A_nhwc, dA_nhwc, out_hw, resize=bilinear_resize):
    return gradcam_tail(np.transpose(A_nhwc, (0, 3, 1, 2)), np.transpose(dA_nhwc, (0, 3, 1, 2)), out_hw, resize)


def jet_lut_bgr() -> np.ndarray:
    """(256,3) uint8 BGR rows = cv2.applyColorMap(arange(256), COLORMAP_JET)."""
    return np.load(os.path.join(_HERE, "..", "tests", "golden", "jet_lut.npy"))


def show_cam_on_image(img_rgb01: np.ndarray, cam: np.ndarray, use_rgb: bool = True,
                      image_weight: float = 0.5) -> np.ndarray:
    """pytorch_grad_cam.utils.image.show_cam_on_image (GRADCAM.py:67)."""
    lut = jet_lut_bgr()
    heat = lut[np.uint8(255 * cam)]                                           # BGR
    if use_rgb:
        heat = heat[..., ::-1]
    heat = np.float32(heat) / 255
    if np.max(img_rgb01) > 1:
        raise Exception("The input image should np.float32 in the range [0, 1]")
    out = (1 - image_weight) * heat + image_weight * img_rgb01
    out = out / np.max(out)
    return np.uint8(255 * out)


def heatmap_u8(cam: np.ndarray) -> np.ndarray:
    """GRADCAM.py:70."""
    return (cam * 255).astype(np.uint8)


def saliency_map(d_input: np.ndarray) -> np.ndarray:
    """explainability.py:72-74 (float stage and the truncating uint8 stage)."""
    s = np.abs(d_input).max(axis=-1)
    s = (s - s.min()) / (s.max() - s.min() + 1e-8)
    return s, np.uint8(s * 255)
