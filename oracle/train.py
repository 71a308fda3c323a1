"""CPU restatement of the training step -- TEST INFRASTRUCTURE.

* mean gradients over a batch: per-sample ``_compute_sample_grads`` summed and divided by n
  (``Classes/CNNModel.py:282-355, 438-464``); top gradient ``probs - y`` (:298-299);
* ``sgd_clip_step``: ``_apply_grads`` + ``_clip_grad`` (:217-222, 372-394): every dW / db / dF / db_conv clipped separately
  to L2 norm 5.0 with ``grad * (max_norm / (norm + 1e-6))``, then ``w -= lr * g``;
* ``adam_step``: ``torch.optim.Adam(model.parameters(), lr)`` exactly as ``ADCNNM.py:88`` constructs it.
Pinned by tests/golden/ref_numpy_train.npz (reference ``_compute_sample_grads`` / ``_apply_grads`` run by make_golden.py).
"""
import numpy as np
import torch

from . import cnn as ocnn


def mean_grads(cfg, params, x, labels, dropout=None, dropout_in_backward=True):
    """{'conv_w': [(F,k,k,C)], 'conv_b': [(F,)], 'dense_w': [(out,in)], 'dense_b': [(out,)]}, loss per sample.

    dropout: per hidden layer [B,units] multipliers; dropout_in_backward=False is the NumPy reference (its backward ignores
    the mask, Classes/CNNModel.py:307-316), True is autograd (ADCNNM.py)."""
    cache = ocnn.forward(cfg, params, x, dropout=dropout)
    labels = np.asarray(labels)
    d_top = ocnn.top_gradient(cache, labels, "softmax_ce")
    _, _, wg = ocnn.backward(cfg, params, cache, d_top, through_input=True, want_wgrads=True,
                              dropout_in_backward=dropout_in_backward)
    out = {"conv_w": [w[0].mean(dim=0).numpy() for w in wg["conv"]], "conv_b": [w[1].mean(dim=0).numpy() for w in wg["conv"]],
           "dense_w": [w[0].mean(dim=0).numpy() for w in wg["dense"]], "dense_b": [w[1].mean(dim=0).numpy() for w in wg["dense"]]}
    p = cache.probs.numpy()
    loss = -np.log(np.clip(p[np.arange(len(labels)), labels], 1e-12, 1.0))          # Classes/CNNModel.py:360-367
    return out, loss


def _clip(g, max_norm):
    n = np.linalg.norm(g)
    return g * (max_norm / (n + 1e-6)) if n > max_norm else g


def sgd_clip_step(params: ocnn.Params, grads, lr, max_norm=5.0) -> ocnn.Params:
    return ocnn.Params([w - lr * _clip(g, max_norm) for w, g in zip(params.conv_w, grads["conv_w"])],
                       [b - lr * _clip(g, max_norm) for b, g in zip(params.conv_b, grads["conv_b"])],
                       [w - lr * _clip(g, max_norm) for w, g in zip(params.dense_w, grads["dense_w"])],
                       [b - lr * _clip(g, max_norm) for b, g in zip(params.dense_b, grads["dense_b"])])


def adam_steps(params: ocnn.Params, grads_seq, lr, betas=(0.9, 0.999), eps=1e-8) -> ocnn.Params:
    """Apply torch.optim.Adam for every gradient set in grads_seq (same tensors, consecutive steps)."""
    flat = [torch.tensor(a, dtype=torch.float64, requires_grad=True)
            for a in list(params.conv_w) + list(params.conv_b) + list(params.dense_w) + list(params.dense_b)]
    opt = torch.optim.Adam(flat, lr=lr, betas=betas, eps=eps)
    for grads in grads_seq:
        gl = list(grads["conv_w"]) + list(grads["conv_b"]) + list(grads["dense_w"]) + list(grads["dense_b"])
        for t, g in zip(flat, gl):
            t.grad = torch.tensor(g, dtype=torch.float64)
        opt.step()
    arrs = [t.detach().numpy() for t in flat]
    nc, nd = len(params.conv_w), len(params.dense_w)
    return ocnn.Params(arrs[:nc], arrs[nc:2 * nc], arrs[2 * nc:2 * nc + nd], arrs[2 * nc + nd:])
