"""Data-parallel training step (SURVEY 8 row f4 / BASELINE config 5).

One process per GPU (torchrun); each rank runs forward + backward on its shard of the global batch in libbcad, the flat
gradient vectors are summed with ONE collective (NCCL all-reduce over NVLink; gloo on CPU in the tests) and divided by the
world size -- "averaged gradient over the global batch" (Classes/CNNModel.py:459-464) -- then every rank applies the same
optimiser step to its replica.  The first dense layer's weight gradient is >99 % of the bytes: it is
reduced as its own bucket, launched first and asynchronously, so it overlaps the small buckets.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from .engine import Engine


def allreduce_mean_(grads: torch.Tensor, big_slice: Optional[slice] = None):
    """In-place mean of `grads` across ranks; `big_slice` (the fc1 weight gradient) goes first as its own async bucket."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return grads
    world = dist.get_world_size()
    handles = []
    if big_slice is not None:
        handles.append(dist.all_reduce(grads[big_slice], op=dist.ReduceOp.SUM, async_op=True))
        if big_slice.start > 0:
            handles.append(dist.all_reduce(grads[:big_slice.start], op=dist.ReduceOp.SUM, async_op=True))
        if big_slice.stop < grads.numel():
            handles.append(dist.all_reduce(grads[big_slice.stop:], op=dist.ReduceOp.SUM, async_op=True))
    else:
        handles.append(dist.all_reduce(grads, op=dist.ReduceOp.SUM, async_op=True))
    for h in handles:
        h.wait()
    grads.div_(world)
    return grads


class DataParallelTrainer:
    def __init__(self, engine: Engine, opt: str = "adam", lr: float = 1e-3, max_norm: float = 5.0, betas=(0.9, 0.999), eps: float = 1e-8):
        if engine.uses_tensor_path:
            raise ValueError("training runs on the fp32 path")
        self.engine, self.opt, self.lr, self.max_norm, self.betas, self.eps = engine, opt, lr, max_norm, betas, eps
        wo, we, _, _ = engine.grad_layout(True, 0)
        self._big = slice(wo, wo + we)
        self._grads = None
        self.last_classes = None

    def step(self, x: torch.Tensor, labels) -> torch.Tensor:
        """One optimiser step on this rank's shard; returns the per-sample losses of the shard (CUDA tensor)."""
        eng = self.engine
        x = eng._as_device_input(x)
        self.last_classes, _, _ = eng.predict(x)                 # forward, activations cached in the handle
        self._grads, loss = eng.train_backward(x, labels, self._grads)
        allreduce_mean_(self._grads, self._big)
        eng.apply_update(self._grads, self.opt, self.lr, self.max_norm, self.betas, self.eps)
        return loss
