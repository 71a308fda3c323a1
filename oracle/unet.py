"""CPU restatement of the tiny U-Net encoder front -- TEST INFRASTRUCTURE.

Follows ``Classes/unet.py``:
* ``conv2d(input, kernel, 'same')`` (:13-30): NHWC cross-correlation, kernel (kh,kw,Cin,F), no bias.  QUIRK kept on purpose:
  the output is allocated at the PADDED size (H+2*pad, W+2*pad) (:19-21) and the windows that would run off the padded
  input are skipped (:26-27), so ``out[:, :H, :W]`` is the true same-convolution and the last 2*pad rows/cols are zero.
* ``max_pool`` (:32-43): 2x2/2, floor dims;  ``relu`` (:53-54).
* ``tiny_unet_numpy`` (:61-73): conv(C->16)+ReLU+pool -> conv(16->32)+ReLU+pool -> conv(32->64)+ReLU; kernels are
  ``np.random.randn`` drawn INSIDE the call, in that order (``draw_kernels`` reproduces the stream after ``np.random.seed``).
* ``average_pool`` (``Classes/ImageSegmentation.py:145-163``): non-overlapping mean, floor dims.
Pinned by tests/golden/ref_unet_*.npz (reference functions exec-loaded by tests/golden/make_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F


def conv2d_same_quirk(x: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    k = kernel.shape[0]
    pad = k // 2
    b, h, w, c = x.shape
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    wt = torch.from_numpy(np.asarray(kernel, dtype=np.float64)).permute(3, 2, 0, 1)        # (kh,kw,C,F)->(F,C,kh,kw)
    y = F.conv2d(xt, wt, padding=pad).permute(0, 2, 3, 1).numpy()                           # true same conv [b,h,w,F]
    out = np.zeros((b, h + 2 * pad, w + 2 * pad, kernel.shape[3]))
    # windows start at i in [0, h+2pad-k]: for odd k that is exactly h positions
    out[:, :y.shape[1], :y.shape[2]] = y
    return out


def max_pool(x: np.ndarray) -> np.ndarray:
    b, h, w, c = x.shape
    h2, w2 = h // 2, w // 2
    return x[:, :2 * h2, :2 * w2].reshape(b, h2, 2, w2, 2, c).max(axis=(2, 4))


def relu(x):
    return np.maximum(0, x)


def draw_kernels(in_channels: int, seed: int):
    """The three kernels tiny_unet_numpy draws after ``np.random.seed(seed)`` (same order, same stream)."""
    rs = np.random.RandomState(seed)
    return [rs.randn(3, 3, in_channels, 16), rs.randn(3, 3, 16, 32), rs.randn(3, 3, 32, 64)]


def tiny_unet(x: np.ndarray, kernels) -> np.ndarray:
    c1 = relu(conv2d_same_quirk(x, kernels[0]))
    p1 = max_pool(c1)
    c2 = relu(conv2d_same_quirk(p1, kernels[1]))
    p2 = max_pool(c2)
    return relu(conv2d_same_quirk(p2, kernels[2]))


def average_pool(x: np.ndarray, pool_size: int = 5) -> np.ndarray:
    b, h, w, c = x.shape
    hn, wn = h // pool_size, w // pool_size
    return x[:, :hn * pool_size, :wn * pool_size].reshape(b, hn, pool_size, wn, pool_size, c).mean(axis=(2, 4))
