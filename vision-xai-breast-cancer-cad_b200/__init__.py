"""B200-native predict + Grad-CAM behind the call surface of ClearanceC137/vision-xai-breast-cancer-cad.

Import as ``bcad_b200`` (the directory name carries the reference's hyphens; ``bcad_b200.py`` at the
repo root registers this directory under that importable name).

Modules mirror the reference files they stand in for:

=====================  ===========================================================
``CNNModel``           Classes/CNNModel.py  (NumPy CNN: ``CNNModel``, ``load_weights``)
``ADCNNM``             WebApplicationPrototype/ADCNNM.py (torch CNN, ``load_trained_model``)
``explainability``     WebApplicationPrototype/explainability.py
``GRADCAM``            WebApplicationPrototype/GRADCAM.py
``ExplainableAI``      Classes/ExplainableAI.py
``unet``               Classes/unet.py layer functions (tiny U-Net encoder front)
``bottleneck``         app.py:466-489 ``process_bottleneck_features`` (CHW -> HWC + cv2 bilinear resize)
``training``           data-parallel training step (wgrad kernels + one gradient all-reduce)
``engine``             batched / sharded driver over the libbcad C-ABI (include/bcad.h)
=====================  ===========================================================

Everything computes inside ``libbcad.so`` (hand-written sm_100a CUDA).  There is no CPU or
PyTorch-op fallback: importing works anywhere, calling needs the built library and a B200.
"""
from . import _lib  # noqa: F401
from .engine import Engine, NetSpec, ShardedEngine, gradcam_tail, gray_preprocess, overlay, shard_bounds  # noqa: F401

__all__ = ["Engine", "NetSpec", "ShardedEngine", "gradcam_tail", "gray_preprocess", "overlay", "shard_bounds"]
