"""CPU restatement of the reference CNN (both flavours) -- TEST INFRASTRUCTURE.

Follows, function by function:

* conv + bias + LeakyReLU, valid cross-correlation, HWC, filters (F,k,k,C):
  ``Classes/CNNModel.py:227-240``  (torch flavour: ``ADCNNM.py:48,76`` --
  ``Conv2d(padding=1)`` + ``F.leaky_relu`` with the DEFAULT slope 0.01).
* 2x2/2 max-pool, floor dims, switches = (patch == max) i.e. every tie marked:
  ``Classes/CNNModel.py:245-261``; backward ``:263-277`` (torch flavour routes
  to the first maximum in row-major window order: ``nn.MaxPool2d`` autograd).
* dense + LeakyReLU, output + softmax(clip +-50, /(sum+1e-12)):
  ``Classes/CNNModel.py:177-196, 203-212``; HWC flatten ``:178`` (torch flavour:
  CHW flatten ``ADCNNM.py:77``, raw logits ``ADCNNM.py:78``).
* explain-backward (activation gradients, d_input, weight grads):
  ``WebApplicationPrototype/explainability.py:13-68``.

Everything is batched and vectorised with torch CPU ops in float64 (or the
dtype asked for), written from the formulas above -- no autograd, so the
tie semantics are explicit.  Pinned to the reference's own outputs by
``tests/test_oracle_golden.py`` (fixtures made by ``tests/golden/make_golden.py``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class NetConfig:
    """One network = one reference constructor call + the flavour switches."""
    input_shape: Tuple[int, int, int]            # (H, W, C)
    num_classes: int
    conv_layers: Sequence[Tuple[int, int]]       # [(filters, ksize)]
    hidden_units: Sequence[int]
    alpha_conv: float = 0.01
    alpha_dense: float = 0.01
    pad: int = 0                                 # 0 = valid (NumPy), 1 = Conv2d(padding=1) (torch)
    flatten: str = "hwc"                         # "hwc" (NumPy) | "chw" (torch)
    pool_ties: str = "dup"                       # "dup" (NumPy) | "first" (torch)
    head: str = "softmax"                        # "softmax" (NumPy) | "logits" (torch)

    @staticmethod
    def numpy_flavour(input_shape, num_classes, conv_layers=((8, 3), (16, 3)),
                      hidden_units=(128, 64), leaky_alpha=0.01) -> "NetConfig":
        # Classes/CNNModel.py:68
        return NetConfig(tuple(input_shape), num_classes, [tuple(c) for c in conv_layers],
                         list(hidden_units), leaky_alpha, leaky_alpha, 0, "hwc", "dup", "softmax")

    @staticmethod
    def torch_flavour(input_shape, num_classes, conv_layers=((32, 3), (64, 3)),
                      hidden_units=(256, 128), leaky_alpha=0.01) -> "NetConfig":
        # ADCNNM.py:35-39 ; conv slope is F.leaky_relu's default (ADCNNM.py:76)
        return NetConfig(tuple(input_shape), num_classes, [tuple(c) for c in conv_layers],
                         list(hidden_units), 0.01, leaky_alpha, 1, "chw", "first", "logits")

    def shapes(self):
        """[(conv_out (h,w,F), pool_out (h,w,F))] per conv block and the flat size."""
        h, w, c = self.input_shape
        out = []
        for f, k in self.conv_layers:
            ch, cw = h + 2 * self.pad - k + 1, w + 2 * self.pad - k + 1
            ph, pw = ch // 2, cw // 2
            out.append(((ch, cw, f), (ph, pw, f)))
            h, w, c = ph, pw, f
        return out, h * w * c


@dataclass
class Params:
    """Weights in the NumPy reference's native layouts.

    conv filters (F,k,k,C) + bias (F,); dense (units, in) + bias (units,), the
    dense input index following ``cfg.flatten`` order.
    """
    conv_w: List[np.ndarray]
    conv_b: List[np.ndarray]
    dense_w: List[np.ndarray]      # hidden layers then the output layer
    dense_b: List[np.ndarray]


def init_params(cfg: NetConfig, seed: int = 7, bias_std: float = 0.0) -> Params:
    """Reference initialisers (He-normal conv ``Classes/CNNModel.py:94``, Glorot-uniform
    dense ``:131-132,146-147``, zero biases) from a seeded Generator."""
    rng = np.random.default_rng(seed)
    conv_w, conv_b, dense_w, dense_b = [], [], [], []
    c = cfg.input_shape[2]
    for f, k in cfg.conv_layers:
        conv_w.append(rng.standard_normal((f, k, k, c)) * np.sqrt(2.0 / (k * k * c)))
        conv_b.append(rng.normal(0, bias_std, f) if bias_std > 0 else np.zeros(f))
        c = f
    _, prev = cfg.shapes()
    for units in list(cfg.hidden_units) + [cfg.num_classes]:
        lim = np.sqrt(6.0 / (prev + units))
        dense_w.append(rng.uniform(-lim, lim, (units, prev)))
        dense_b.append(rng.normal(0, bias_std, units) if bias_std > 0 else np.zeros(units))
        prev = units
    return Params(conv_w, conv_b, dense_w, dense_b)


def synth_images(n: int, shape=(256, 256, 1), seed: int = 20251018, kind: str = "gauss") -> np.ndarray:
    """Seeded synthetic inputs: per-image standardised float32 [n,H,W,C].

    kind="mammo": >=40 % exact-zero background + smooth blob, to exercise pool ties.
    """
    rng = np.random.default_rng(seed)
    h, w, c = shape
    x = rng.standard_normal((n, h, w, c)).astype(np.float32)
    if kind == "mammo":
        yy, xx = np.mgrid[0:h, 0:w]
        for i in range(n):
            cy, cx = rng.uniform(0.3, 0.7) * h, rng.uniform(0.1, 0.5) * w
            r = rng.uniform(0.25, 0.4) * min(h, w)
            blob = np.exp(-(((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r)))
            mask = (blob > 0.55).astype(np.float32)
            x[i] = (np.abs(x[i]) * 0.2 + blob[..., None]) * mask[..., None]
        return x.astype(np.float32)          # zeros stay exact zeros (no standardisation)
    mean = x.mean(axis=(1, 2, 3), keepdims=True)
    std = x.std(axis=(1, 2, 3), keepdims=True)
    return ((x - mean) / std).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# layout converters (NumPy reference layout <-> torch reference layout), SURVEY section 7 hard part 1
# ----------------------------------------------------------------------------------------------
def fc1_hwc_to_chw(w: np.ndarray, h: int, wd: int, c: int) -> np.ndarray:
    u = w.shape[0]
    return w.reshape(u, h, wd, c).transpose(0, 3, 1, 2).reshape(u, -1)


def fc1_chw_to_hwc(w: np.ndarray, h: int, wd: int, c: int) -> np.ndarray:
    u = w.shape[0]
    return w.reshape(u, c, h, wd).transpose(0, 2, 3, 1).reshape(u, -1)


def params_to_state_dict(cfg: NetConfig, p: Params) -> Dict[str, torch.Tensor]:
    """NumPy-layout Params (dense[0] columns in cfg.flatten order) -> ADCNNM state_dict
    (keys ``convs.{i}.weight/bias``, ``fc.{0,3,6,..}.weight/bias``: ADCNNM.py:59-70)."""
    sd = {}
    for i, (w, b) in enumerate(zip(p.conv_w, p.conv_b)):
        sd[f"convs.{i}.weight"] = torch.tensor(np.ascontiguousarray(w.transpose(0, 3, 1, 2)), dtype=torch.float32)
        sd[f"convs.{i}.bias"] = torch.tensor(b, dtype=torch.float32)
    shapes, _ = cfg.shapes()
    ph, pw, pc = shapes[-1][1]
    for j, (w, b) in enumerate(zip(p.dense_w, p.dense_b)):
        if j == 0 and cfg.flatten == "hwc":
            w = fc1_hwc_to_chw(w, ph, pw, pc)
        sd[f"fc.{3 * j}.weight"] = torch.tensor(np.ascontiguousarray(w), dtype=torch.float32)
        sd[f"fc.{3 * j}.bias"] = torch.tensor(b, dtype=torch.float32)
    return sd


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------
def _leaky(z, a):
    return torch.where(z > 0, z, a * z)           # strict ">" : Classes/CNNModel.py:184,239


def _dleaky(z, a):
    one = torch.ones((), dtype=z.dtype)
    return torch.where(z > 0, one, one * a)       # built in z's dtype (a float32 0.01 would cost 2e-9)


def softmax_clip(z: torch.Tensor) -> torch.Tensor:
    """Classes/CNNModel.py:203-212 (rows of z)."""
    z = z.double().clamp(-50.0, 50.0)
    z = z - z.max(dim=-1, keepdim=True).values
    e = torch.exp(z)
    s = e.sum(dim=-1, keepdim=True)
    out = e / (s + 1e-12)
    uniform = torch.full_like(out, 1.0 / z.shape[-1])
    return torch.where(s == 0, uniform, out)


@dataclass
class Cache:
    cfg: NetConfig
    x: torch.Tensor                                   # [B,H,W,C]
    conv_in: List[torch.Tensor] = field(default_factory=list)    # NHWC
    conv_out: List[torch.Tensor] = field(default_factory=list)   # NHWC post-LeakyReLU
    pool_out: List[torch.Tensor] = field(default_factory=list)   # NHWC
    switches: List[torch.Tensor] = field(default_factory=list)   # NHWC float mask (per cfg.pool_ties)
    dense_in: List[torch.Tensor] = field(default_factory=list)   # [B,in]
    z: List[torch.Tensor] = field(default_factory=list)          # pre-activations, last = logits
    logits: Optional[torch.Tensor] = None
    probs: Optional[torch.Tensor] = None
    dropout: Optional[list] = None                               # per hidden layer [B,units] multipliers (training)


def _pool_with_switches(y_nchw: torch.Tensor, ties: str):
    b, c, h, w = y_nchw.shape
    h2, w2 = h // 2, w // 2
    win = y_nchw[:, :, :2 * h2, :2 * w2].reshape(b, c, h2, 2, w2, 2)
    p = win.amax(dim=(3, 5))
    eq = (win == p[:, :, :, None, :, None])
    if ties == "first":
        # first maximum in row-major window order (0,0),(0,1),(1,0),(1,1)
        flat = eq.permute(0, 1, 2, 4, 3, 5).reshape(b, c, h2, w2, 4)
        first = flat.to(torch.int8).argmax(dim=-1)
        oh = F.one_hot(first, 4).bool().reshape(b, c, h2, w2, 2, 2).permute(0, 1, 2, 4, 3, 5)
        eq = oh
    sw = torch.zeros_like(y_nchw)
    sw[:, :, :2 * h2, :2 * w2] = eq.reshape(b, c, 2 * h2, 2 * w2).to(y_nchw.dtype)
    return p, sw


def forward(cfg: NetConfig, params: Params, x, dtype=torch.float64, dropout=None) -> Cache:
    """Batched forward with the reference's caches (Classes/CNNModel.py:162-198).

    dropout: training mode only -- one [B,units] multiplier array (0 or 1/(1-rate)) per hidden dense layer, applied after
    the activation (Classes/CNNModel.py:186-188 / nn.Dropout)."""
    x = torch.as_tensor(np.asarray(x)).to(dtype)
    if x.dim() == 3:
        x = x[None]
    cache = Cache(cfg, x)
    cur = x.permute(0, 3, 1, 2)                        # NCHW for F.conv2d
    for (w, b) in zip(params.conv_w, params.conv_b):
        wt = torch.as_tensor(w).to(dtype).permute(0, 3, 1, 2)       # (F,k,k,C)->(F,C,k,k)
        cache.conv_in.append(cur.permute(0, 2, 3, 1))
        y = _leaky(F.conv2d(cur, wt, torch.as_tensor(b).to(dtype), padding=cfg.pad), cfg.alpha_conv)
        p, sw = _pool_with_switches(y, cfg.pool_ties)
        cache.conv_out.append(y.permute(0, 2, 3, 1))
        cache.pool_out.append(p.permute(0, 2, 3, 1))
        cache.switches.append(sw.permute(0, 2, 3, 1))
        cur = p
    bsz = x.shape[0]
    flat = (cur.permute(0, 2, 3, 1) if cfg.flatten == "hwc" else cur).reshape(bsz, -1)
    n_dense = len(params.dense_w)
    for j, (w, b) in enumerate(zip(params.dense_w, params.dense_b)):
        cache.dense_in.append(flat)
        z = flat @ torch.as_tensor(w).to(dtype).T + torch.as_tensor(b).to(dtype)
        cache.z.append(z)
        flat = _leaky(z, cfg.alpha_dense) if j < n_dense - 1 else z
        if dropout is not None and j < n_dense - 1:
            flat = flat * torch.as_tensor(dropout[j]).to(dtype)
    cache.dropout = dropout
    cache.logits = cache.z[-1]
    cache.probs = softmax_clip(cache.logits) if cfg.head == "softmax" else torch.softmax(cache.logits, dim=-1)
    return cache


def predict(cfg: NetConfig, params: Params, x, dtype=torch.float64):
    """Classes/CNNModel.py:524-526 batched: (argmax(probs) -> first max, probs)."""
    c = forward(cfg, params, x, dtype)
    score = c.probs if cfg.head == "softmax" else c.logits      # app.py:589 takes torch.max of the logits
    return score.argmax(dim=-1).numpy(), c.probs.numpy(), c.logits.numpy()


# ----------------------------------------------------------------------------------------------
# explain-backward
# ----------------------------------------------------------------------------------------------
def top_gradient(cache: Cache, class_idx, mode: str) -> torch.Tensor:
    """mode "softmax_ce": probs - onehot (explainability.py:21-22);
    mode "logit": e_c, d(logit_c)/d(logits)  (GRADCAM.py:64 ClassifierOutputTarget)."""
    b, nc = cache.logits.shape
    idx = torch.as_tensor(np.broadcast_to(np.asarray(class_idx), (b,)).copy()).long()
    onehot = F.one_hot(idx, nc).to(cache.logits.dtype)
    if mode == "logit":
        return onehot
    if mode == "softmax_ce":
        return cache.probs.to(cache.logits.dtype) - onehot     # the model's own probabilities (per head)
    raise ValueError(mode)


def backward(cfg: NetConfig, params: Params, cache: Cache, d_top: torch.Tensor,
             through_input: bool = True, want_wgrads: bool = False, dropout_in_backward: bool = True):
    """explainability.py:13-68 batched.

    Returns (conv_act_grads {conv_block: [B,h,w,F]}, d_input [B,H,W,C] or None, wgrads or None).
    conv_act_grads[i] is the gradient w.r.t. the POST-LeakyReLU output of conv block i
    (explainability.py:64 copies d_out before the activation mask is applied).
    """
    dtype = d_top.dtype
    d = d_top
    n_dense = len(params.dense_w)
    wgrads = {"dense": [None] * n_dense, "conv": [None] * len(params.conv_w)} if want_wgrads else None
    for j in reversed(range(n_dense)):
        w = torch.as_tensor(params.dense_w[j]).to(dtype)
        if j < n_dense - 1:
            # autograd masks the gradient with the dropout multipliers; the NumPy reference's _compute_sample_grads does
            # NOT (Classes/CNNModel.py:307-316 uses d_out as is) -- dropout_in_backward=False restates that
            if cache.dropout is not None and dropout_in_backward:
                d = d * torch.as_tensor(cache.dropout[j]).to(dtype)
            d = d * _dleaky(cache.z[j], cfg.alpha_dense)                            # :28-29
        if want_wgrads:
            wgrads["dense"][j] = (d[:, :, None] * cache.dense_in[j][:, None, :], d.clone())   # per-sample outer
        d = d @ w                                                                   # W^T dz  (:26,:34)
    shapes, _ = cfg.shapes()
    ph, pw, pc = shapes[-1][1]
    bsz = d.shape[0]
    g = d.reshape(bsz, ph, pw, pc) if cfg.flatten == "hwc" else d.reshape(bsz, pc, ph, pw).permute(0, 2, 3, 1)
    conv_act_grads = {}
    d_input = None
    for i in reversed(range(len(params.conv_w))):
        y = cache.conv_out[i]
        _, h, w_, f = y.shape
        h2, w2 = h // 2, w_ // 2
        # un-pool: dX[window] += d_out * switches[window]  (Classes/CNNModel.py:263-277)
        up = torch.zeros_like(y)
        up[:, :2 * h2, :2 * w2, :] = g.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
        dA = up * cache.switches[i]
        conv_act_grads[i] = dA
        if i == 0 and not through_input:
            break
        dz = dA * _dleaky(y, cfg.alpha_conv)                                         # :55 (mask on the OUTPUT)
        wt = torch.as_tensor(params.conv_w[i]).to(dtype).permute(0, 3, 1, 2)
        if want_wgrads:
            xin = cache.conv_in[i].permute(0, 3, 1, 2)
            k = wt.shape[-1]
            xp = F.pad(xin, (cfg.pad,) * 4)
            patches = xp.unfold(2, k, 1).unfold(3, k, 1)                            # B,C,h,w,k,k
            dF = torch.einsum("bhwf,bchwuv->bfuvc", dz, patches)
            wgrads["conv"][i] = (dF, dz.sum(dim=(1, 2)))
        g = F.conv_transpose2d(dz.permute(0, 3, 1, 2), wt, padding=cfg.pad).permute(0, 2, 3, 1)  # :60
        if i == 0:
            d_input = g
    return conv_act_grads, d_input, wgrads
