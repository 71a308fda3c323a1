"""Drop-in for ``process_bottleneck_features`` of ``WebApplicationPrototype/app.py:466-489``: the producer that feeds the
basic classifier -- a [C,H,W] bottleneck feature map becomes HWC and is resized with ``cv2.resize(INTER_LINEAR)`` to
``resize_shape`` (default 32x32).  Runs in libbcad (``bcad_bottleneck_resize``), bit-exact with OpenCV's float paths.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _ptr(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


def resize_batch(feats, resize_shape=(32, 32), layout: str = "chw", device: int = 0) -> torch.Tensor:
    """feats [B,C,H,W] (layout "chw") or [B,H,W,C] ("hwc"), tensor or ndarray -> CUDA tensor fp32 [B,h,w,C].
    ``resize_shape`` is cv2's dsize = (width, height)."""
    if not torch.cuda.is_available():
        raise RuntimeError("bcad_b200 needs a CUDA device (no CPU fallback)")
    lib = _lib.load()
    dev = torch.device("cuda", device)
    x = torch.as_tensor(feats).to(device=dev, dtype=torch.float32).contiguous()
    if x.dim() != 4:
        raise ValueError("expected a 4-D batch")
    if layout == "chw":
        B, Cc, H, W = x.shape
    elif layout == "hwc":
        B, H, W, Cc = x.shape
    else:
        raise ValueError("layout must be 'chw' or 'hwc'")
    ow, oh = int(resize_shape[0]), int(resize_shape[1])
    with torch.cuda.device(dev):
        out = torch.empty((B, oh, ow, Cc), device=dev, dtype=torch.float32)
        _lib.check(lib.bcad_bottleneck_resize(_ptr(x), B, Cc, H, W, 0 if layout == "chw" else 1, oh, ow, _ptr(out),
                                              C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def process_bottleneck_features(feat, resize_shape=(32, 32)):
    """Same signature and return as the reference: one feature map -> np.ndarray [h, w, C] float32.
    A tensor is taken as [C,H,W]; an ndarray is transposed when ``shape[0] < shape[2]`` (app.py:480-484)."""
    if isinstance(feat, torch.Tensor):
        layout = "chw"
    else:
        feat = np.asarray(feat)
        layout = "chw" if feat.shape[0] < feat.shape[2] else "hwc"
    return resize_batch(torch.as_tensor(feat)[None], resize_shape, layout)[0].cpu().numpy()
