"""Drop-in for the layer functions of ``Classes/unet.py`` (tiny U-Net encoder front).

Same names and NHWC semantics as the reference -- ``conv2d(input, kernel, padding='same')`` INCLUDING its quirk
(output allocated at the padded size, trailing rows/cols left zero: unet.py:19-27), ``max_pool``, ``relu``,
``tiny_unet_numpy(input_image)`` (kernels drawn with ``np.random.randn`` inside the call, in the reference's order) and
``average_pool`` of ``Classes/ImageSegmentation.py:145-163`` -- computed by libbcad's conv / pool kernels.
The reference's import-time script (unet.py:75-114, needs a missing ``preprocessing`` module) is not reproduced.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _dev(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def conv_block(x, kernel, alpha=1.0, want_y=True, want_pool=False, padded_output=True, bias=None):
    """conv2d (+ LeakyReLU_alpha, alpha=1 identity / 0 ReLU) (+ 2x2 max-pool) on CUDA tensors; returns (y, pooled)."""
    lib = _lib.load()
    x, kernel = _dev(x), _dev(kernel)
    B, H, W, Cin = x.shape
    k, k2, cin2, F = kernel.shape
    if k != k2 or cin2 != Cin:
        raise ValueError(f"kernel shape {tuple(kernel.shape)} does not match input channels {Cin}")
    pad = k // 2
    Ho, Wo = (H + 2 * pad, W + 2 * pad) if padded_output else (H + 2 * pad - k + 1, W + 2 * pad - k + 1)
    y = torch.empty((B, Ho, Wo, F), device=x.device, dtype=torch.float32) if want_y else None
    p = torch.empty((B, Ho // 2, Wo // 2, F), device=x.device, dtype=torch.float32) if want_pool else None
    b = _dev(bias) if bias is not None else None
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    with torch.cuda.device(x.device):
        _lib.check(lib.bcad_conv_block(ptr(x), B, H, W, Cin, ptr(kernel), ptr(b), k, F, pad, float(alpha),
                                       1 if padded_output else 0, ptr(y), ptr(p), _stream(x)))
    return y, p


def conv2d(input, kernel, padding="same"):
    """Classes/unet.py:13-30 (NumPy in, NumPy float64 out)."""
    if padding != "same":
        raise ValueError("the reference only ever calls conv2d(..., 'same')")
    y, _ = conv_block(input, kernel, alpha=1.0, want_y=True)
    return y.double().cpu().numpy()


def max_pool(input):
    """Classes/unet.py:32-43: 2x2/2 max, floor dims (an identity 1x1 convolution with the pool fused)."""
    x = _dev(input)
    c = x.shape[-1]
    eye = torch.eye(c, device=x.device).reshape(1, 1, c, c)
    _, p = conv_block(x, eye, alpha=1.0, want_y=False, want_pool=True, padded_output=False)
    return p.double().cpu().numpy()


def relu(x):
    return np.maximum(0, x)


def tiny_unet(input_image, kernels, as_numpy=True):
    """conv(C->16)+ReLU+pool -> conv(16->32)+ReLU+pool -> conv(32->64)+ReLU with explicit kernels (kh,kw,Cin,F);
    ReLU and the pools are fused into the conv launches (3 launches)."""
    _, p1 = conv_block(input_image, kernels[0], alpha=0.0, want_y=False, want_pool=True)
    _, p2 = conv_block(p1, kernels[1], alpha=0.0, want_y=False, want_pool=True)
    bn, _ = conv_block(p2, kernels[2], alpha=0.0, want_y=True)
    return bn.double().cpu().numpy() if as_numpy else bn


def tiny_unet_numpy(input_image):
    """Classes/unet.py:61-73: kernels come from the global NumPy stream, drawn in the reference's order."""
    c = np.asarray(input_image).shape[-1]
    kernels = [np.random.randn(3, 3, c, 16), np.random.randn(3, 3, 16, 32), np.random.randn(3, 3, 32, 64)]
    return tiny_unet(input_image, kernels)


def average_pool(input, pool_size=5, as_numpy=True):
    """Classes/ImageSegmentation.py:145-163."""
    lib = _lib.load()
    x = _dev(input)
    B, H, W, Cc = x.shape
    out = torch.empty((B, H // pool_size, W // pool_size, Cc), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.bcad_avg_pool(C.c_void_p(x.data_ptr()), B, H, W, Cc, int(pool_size), C.c_void_p(out.data_ptr()), _stream(x)))
    return out.double().cpu().numpy() if as_numpy else out
