"""The reference's practical CPU path as a timed baseline -- TEST / BENCH INFRASTRUCTURE.

Port ("kind": "port") of what the reference runs on CPU for predict + Grad-CAM:

* the torch CNN of ``WebApplicationPrototype/ADCNNM.py:34-78`` restated module for module
  (Conv2d(padding=1) -> F.leaky_relu (default slope) -> MaxPool2d(2); reshape CHW; Linear/LeakyReLU/Dropout),
  run in eval mode on all host threads (oneDNN), fp32;
* Grad-CAM as ``pytorch_grad_cam.GradCAM`` does it at ``GRADCAM.py:53,64``: hook the target activation
  (last conv block, post-LeakyReLU), back-propagate the summed class logits with autograd, then the NumPy tail
  (``oracle.gradcam.gradcam_tail``; cv2.resize when OpenCV is importable, its NumPy twin otherwise).

The reference sources cannot travel to the GPU box (Python, not copied), so this port is what ``bench.py`` times
there.  ``tests/test_oracle_golden.py`` pins it to the reference's own ADCNNM outputs through the shared
state_dict layout.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import cnn as ocnn
from . import gradcam as ogc


class TorchCNN(nn.Module):
    def __init__(self, input_shape, num_classes, conv_layers=((32, 3), (64, 3)), hidden_units=(256, 128),
                 dropout_rate=0.3, leaky_alpha=0.01):
        super().__init__()
        H, W, C = input_shape
        self.convs = nn.ModuleList()
        self.pools = nn.ModuleList()
        cin = C
        for cout, k in conv_layers:
            self.convs.append(nn.Conv2d(cin, cout, k, padding=1))
            self.pools.append(nn.MaxPool2d(2))
            cin = cout
        with torch.no_grad():
            d = torch.zeros(1, C, H, W)
            for conv, pool in zip(self.convs, self.pools):
                d = pool(F.leaky_relu(conv(d), negative_slope=leaky_alpha))
            flat = d.view(1, -1).size(1)
        layers, prev = [], flat
        for u in hidden_units:
            layers += [nn.Linear(prev, u), nn.LeakyReLU(leaky_alpha), nn.Dropout(dropout_rate)]
            prev = u
        layers.append(nn.Linear(prev, num_classes))
        self.fc = nn.Sequential(*layers)

    def forward(self, x, keep=None):
        x = x.permute(0, 3, 1, 2)
        for i, (conv, pool) in enumerate(zip(self.convs, self.pools)):
            a = F.leaky_relu(conv(x))
            if keep is not None and i == len(self.convs) - 1:
                a.retain_grad()
                keep.append(a)
            x = pool(a)
        return self.fc(x.reshape(x.size(0), -1))


def build(cfg: ocnn.NetConfig, params: ocnn.Params) -> TorchCNN:
    m = TorchCNN(cfg.input_shape, cfg.num_classes, cfg.conv_layers, cfg.hidden_units, 0.3, cfg.alpha_dense)
    m.load_state_dict(ocnn.params_to_state_dict(cfg, params))
    return m.eval()


def _resize():
    try:
        import cv2
        return lambda img, H, W: cv2.resize(img, (W, H))
    except Exception:       # pragma: no cover
        return ogc.bilinear_resize


def predict_gradcam(model: TorchCNN, x: torch.Tensor, class_idx=None):
    """x: fp32 [B,H,W,C] -> (classes [B], logits [B,nc], heatmaps float32 [B,H,W])."""
    keep = []
    logits = model(x, keep)
    cls = logits.argmax(dim=1)
    tgt = cls if class_idx is None else torch.as_tensor(class_idx).long()
    model.zero_grad(set_to_none=True)
    logits.gather(1, tgt[:, None]).sum().backward()
    A = keep[0].detach().numpy()
    dA = keep[0].grad.numpy()
    heat = ogc.gradcam_tail(A, dA, x.shape[1:3], resize=_resize())
    return cls.numpy(), logits.detach().numpy(), heat


def time_predict_gradcam(cfg, params, n_images: int, batch: int = 32, repeats: int = 1, seed: int = 20251018):
    """images/s of predict + Grad-CAM over `n_images` synthetic inputs in batches of `batch` (1 warm-up batch)."""
    model = build(cfg, params)
    x = torch.from_numpy(ocnn.synth_images(n_images, cfg.input_shape, seed=seed))
    predict_gradcam(model, x[:min(batch, n_images)])
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        for s in range(0, n_images, batch):
            predict_gradcam(model, x[s:s + batch])
        best = min(best, time.perf_counter() - t0)
    return n_images / best, best
