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


class UnetFront:
    """``tiny_unet`` + ``average_pool`` as ONE tensor-core pipeline (libbcad ``bcad_unet_*``: conv2 / conv3 as tcgen05 implicit
    GEMMs, fp16 operands, fp32 accumulation) for batches of single-channel images whose H and W are multiples of 4 -- BASELINE
    config 3's front at batch 256.  Results agree with the reference functions to the 16-bit mode's tolerance (1e-2 of the
    map's scale); ``tiny_unet`` / ``average_pool`` above are the fp32 route for every other shape."""

    def __init__(self, H, W, kernels=None, max_batch=256, device=0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("libbcad needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.H, self.W, self.device = int(H), int(W), int(device)
        self.tdev = torch.device("cuda", self.device)
        self._h = C.c_void_p()
        _lib.check(self.lib.bcad_unet_create(self.H, self.W, int(max_batch), self.device, C.byref(self._h)))
        if kernels is not None:
            self.set_kernels(kernels)

    def set_kernels(self, kernels):
        """kernels: the three arrays ``tiny_unet_numpy`` draws -- (3,3,1,16), (3,3,16,32), (3,3,32,64)."""
        ks = [np.ascontiguousarray(k, dtype=np.float32) for k in kernels]
        if [k.shape for k in ks] != [(3, 3, 1, 16), (3, 3, 16, 32), (3, 3, 32, 64)]:
            raise ValueError(f"kernel shapes {[k.shape for k in ks]} are not the tiny U-Net's")
        _lib.check(self.lib.bcad_unet_set_kernels(self._h, *[C.c_void_p(k.ctypes.data) for k in ks]))

    def out_shape(self, avg_pool=3):
        h, w, c = C.c_int(), C.c_int(), C.c_int()
        _lib.check(self.lib.bcad_unet_out_shape(self._h, int(avg_pool), C.byref(h), C.byref(w), C.byref(c)))
        return h.value, w.value, c.value

    def forward(self, x, avg_pool=3, out=None):
        """x: [B,H,W] or [B,H,W,1] (CUDA tensor or array) -> fp32 CUDA tensor [B,h,w,64]: ``average_pool(tiny_unet(x), avg_pool)``,
        or the ``bn`` tensor itself with ``avg_pool=0``."""
        x = x.to(device=self.tdev, dtype=torch.float32) if isinstance(x, torch.Tensor) else \
            torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(self.tdev)
        if x.dim() == 4 and x.shape[-1] == 1:
            x = x[..., 0]
        if x.dim() != 3 or tuple(x.shape[1:]) != (self.H, self.W):
            raise ValueError(f"input shape {tuple(x.shape)} does not match [B,{self.H},{self.W}(,1)]")
        x = x.contiguous()
        h, w, c = self.out_shape(avg_pool)
        if out is None:
            out = torch.empty((x.shape[0], h, w, c), device=self.tdev, dtype=torch.float32)
        with torch.cuda.device(self.tdev):
            _lib.check(self.lib.bcad_unet_forward(self._h, C.c_void_p(x.data_ptr()), int(x.shape[0]), int(avg_pool),
                                                  C.c_void_p(out.data_ptr()), _stream(x)))
        return out

    @property
    def launch_count(self):
        return int(self.lib.bcad_unet_launch_count(self._h))

    def set_profiling(self, on: bool):
        _lib.check(self.lib.bcad_unet_set_profiling(self._h, 1 if on else 0))

    def last_profile(self):
        """[(stage, ms)] of the last forward's last chunk (needs set_profiling(True) before it)."""
        out = []
        for i in range(5):
            buf, ms = C.create_string_buffer(64), C.c_float()
            _lib.check(self.lib.bcad_unet_profile_get(self._h, i, buf, 64, C.byref(ms)))
            out.append((buf.value.decode(), float(ms.value)))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.bcad_unet_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
