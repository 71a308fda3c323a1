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


def allreduce_sum_(grads: torch.Tensor, big_slice: Optional[slice] = None):
    """In-place SUM of `grads` across ranks (bucketed like allreduce_mean_)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return grads
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
    return grads


class DataParallelTrainer:
    """``overlap=True`` (default under an initialised process group): the backward runs in two parts (bcad_train_backward_part) --
    the dense layers first, whose gradients (fc1: 99.9 % of the bytes) go into an asynchronous NCCL all-reduce at once, then the
    conv blocks, which run WHILE that all-reduce is on the wire; the small conv bucket follows.  ``equal_shards=False`` weights
    every rank's mean gradient by its shard size (the mean over the global batch, Classes/CNNModel.py:459-464, also when the
    last shards are smaller); with equal shards the plain average is that mean already."""

    def __init__(self, engine: Engine, opt: str = "adam", lr: float = 1e-3, max_norm: float = 5.0, betas=(0.9, 0.999), eps: float = 1e-8,
                 overlap: bool = True, equal_shards: bool = True):
        if engine.uses_tensor_path:
            raise ValueError("training runs on the fp32 path")
        self.engine, self.opt, self.lr, self.max_norm, self.betas, self.eps = engine, opt, lr, max_norm, betas, eps
        self.overlap, self.equal_shards = overlap, equal_shards
        wo, we, _, _ = engine.grad_layout(True, 0)
        self._big = slice(wo, wo + we)
        self._dense0 = wo                    # the flat vector is [conv tensors | dense tensors]: everything from here on is dense
        self._grads = None
        self.last_classes = None
        self.last_comm_ms = None

    def step(self, x: torch.Tensor, labels) -> torch.Tensor:
        """One optimiser step on this rank's shard; returns the per-sample losses of the shard (CUDA tensor)."""
        eng = self.engine
        x = eng._as_device_input(x)
        self.last_classes, _, _ = eng.predict(x)                 # forward, activations cached in the handle
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        if world == 1:
            self._grads, loss = eng.train_backward(x, labels, self._grads)
            eng.apply_update(self._grads, self.opt, self.lr, self.max_norm, self.betas, self.eps)
            return loss
        n_local = int(x.shape[0])
        pre, post = None, float(world)                           # equal shards: sum of the ranks' means / world
        if not self.equal_shards:                                # else: sum_r n_r * mean_r / sum_r n_r
            cnt = torch.tensor([float(n_local)], device=eng.tdev)
            dist.all_reduce(cnt)
            pre, post = float(n_local), float(cnt.item())
        if not self.overlap:
            self._grads, loss = eng.train_backward(x, labels, self._grads)
            if pre is not None:
                self._grads.mul_(pre)
            allreduce_sum_(self._grads, self._big)
            self._grads.div_(post)
            eng.apply_update(self._grads, self.opt, self.lr, self.max_norm, self.betas, self.eps)
            return loss
        # part 1: dense layers; their all-reduce starts as soon as the kernels are queued (NCCL's stream waits for them on the device)
        self._grads, loss = eng.train_backward(x, labels, self._grads, part=1)
        g = self._grads
        if pre is not None:
            g[self._dense0:].mul_(pre)
        handles = [dist.all_reduce(g[self._big], op=dist.ReduceOp.SUM, async_op=True)]
        if self._big.stop < g.numel():
            handles.append(dist.all_reduce(g[self._big.stop:], op=dist.ReduceOp.SUM, async_op=True))
        # part 2: conv blocks, overlapping the fc1 bucket on the wire
        eng.train_backward(x, labels, g, part=2, loss=loss)
        if pre is not None:
            g[:self._dense0].mul_(pre)
        handles.append(dist.all_reduce(g[:self._dense0], op=dist.ReduceOp.SUM, async_op=True))
        for h in handles:
            h.wait()
        g.div_(post)
        eng.apply_update(g, self.opt, self.lr, self.max_norm, self.betas, self.eps)
        return loss
