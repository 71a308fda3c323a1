"""``Classes/ExplainableAI.py`` filled in: the reference declares the API and leaves every method ``pass``
(Classes/ExplainableAI.py:8-16).  Same attributes and signatures; the work is done by libbcad."""
from __future__ import annotations

import numpy as np
import torch

from . import engine as _engine


class ExplainableAI:
    def __init__(self):
        self.heatmap = None                   # Grad-CAM heatmap highlighting important image regions
        self.last_conv_layer = None           # Last convolutional layer used for Grad-CAM
        self.colormap = "jet"                 # Color map used for visualizing heatmap (default: 'jet')

    def generate_heatmap(self, model, image, class_index, grad_mode="logit"):
        """image: (H,W,C) model input (or [B,H,W,C]); -> float32 heatmap (H,W) (or [B,H,W]) in [0,1]."""
        eng = getattr(model, "fast_engine", None) or model.engine
        x = np.asarray(image, dtype=np.float32)
        single = x.ndim == 3
        xb = x[None] if single else x
        _, _, _, heat = eng.predict_explain(xb, class_index, grad_mode)
        self.last_conv_layer = len(eng.spec.conv_layers) - 1
        out = heat.cpu().numpy()
        self.heatmap = out[0] if single else out
        return self.heatmap

    def overlay_heatmap(self, image, heatmap):
        """image: grayscale (H,W) in [0,1] (or 0-255); -> uint8 RGB overlay (H,W,3) (show_cam_on_image)."""
        if self.colormap != "jet":
            raise ValueError("only the 'jet' colormap of the reference is built")
        img = np.asarray(image, dtype=np.float32)
        if img.ndim == 3:
            img = img[..., 0]
        if img.max() > 1:
            img = img / 255.0
        dev = torch.device("cuda")
        ov, _ = _engine.overlay(torch.from_numpy(img)[None].to(dev), torch.from_numpy(np.asarray(heatmap, np.float32))[None].to(dev),
                                want_heat_u8=False)
        return ov[0].cpu().numpy()

    def visualize_prediction(self, image, heatmap):
        """Returns the overlay (the reference would display it; no GUI dependency here)."""
        return self.overlay_heatmap(image, heatmap)
