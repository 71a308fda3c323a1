"""Every-image comparison of a device engine with the float64 oracle -- TEST INFRASTRUCTURE (tests/ and the untimed `check`
leg of bench.py; the product path never imports this).

The engine object only needs ``max_batch``, ``predict_explain(x, class_idx, grad_mode) -> (cls, probs, logits, heat)`` and
``get_tensor(kind, index, B)`` returning torch tensors (the C-ABI wrapper ``bcad_b200.Engine``)."""
import numpy as np

from . import cnn as ocnn

T_DENSE_Z = 2          # include/bcad.h BCAD_T_DENSE_Z


def compare_all_images(cfg, params, x, eng, modes, tau, batch=32):
    """Run `eng` over all images of `x` and compare every one of them with the float64 oracle -- no image is excluded.

    Grad-CAM is discontinuous where a hidden dense unit's pre-activation z crosses 0 (LeakyReLU' jumps from alpha to 1), so any
    finite-precision forward -- the reference's own float32 one included -- can legitimately pick the other branch for a unit
    whose |z| is below its rounding error.  The oracle heat-map of an image is therefore evaluated with the DEVICE's branch
    for exactly those hidden units whose oracle |z| <= tau (tau = the stated bound on the path's pre-activation error); a unit
    whose branch differs although |z_oracle| > tau is a hard violation and is counted (`mask_violations`, must be 0).

    eng: one engine, or a list [(engine, tau), ...] sharing one oracle sweep (then `tau` is ignored and a list of results, one
    per engine, comes back).  modes: [(class_idx or None, grad_mode)].  Returns one dict per mode with per-image arrays:
      cls_equal [B] bool, logit_err [B], heat_err [B], overridden [B] bool (image needed a branch override), margin [B]
      (oracle top-2 logit gap), logit_absmax [B], plus mask_violations (int)."""
    import torch
    from . import gradcam as ogc
    single = not isinstance(eng, (list, tuple))
    engs = [(eng, tau)] if single else list(eng)
    B = x.shape[0]
    n_hidden = len(cfg.hidden_units)
    last = len(cfg.conv_layers) - 1
    out = [[dict(cls_equal=np.zeros(B, bool), logit_err=np.zeros(B), heat_err=np.zeros(B), overridden=np.zeros(B, bool),
                 margin=np.zeros(B), logit_absmax=np.zeros(B), mask_violations=0) for _ in modes] for _ in engs]
    mb = min(e.max_batch for e, _ in engs)
    for s0 in range(0, B, mb):
        s1 = min(B, s0 + mb)
        dev = []
        for e, _ in engs:
            per_mode = []
            for class_idx, mode in modes:
                ci = None if class_idx is None else np.asarray(class_idx)[s0:s1]
                cls, probs, logits, heat = e.predict_explain(x[s0:s1], ci, mode)
                zs = [e.get_tensor(T_DENSE_Z, j, s1 - s0).cpu().numpy() for j in range(n_hidden)]
                per_mode.append((cls.cpu().numpy(), logits.cpu().numpy(), heat.cpu().numpy(), zs))
            dev.append(per_mode)
        for b0 in range(s0, s1, batch):
            b1 = min(s1, b0 + batch)
            cache = ocnn.forward(cfg, params, x[b0:b1])
            score = cache.probs if cfg.head == "softmax" else cache.logits
            o_cls = score.argmax(dim=-1).numpy()
            lg = cache.logits.numpy()
            srt = np.sort(lg, axis=1)
            margin = srt[:, -1] - srt[:, -2] if lg.shape[1] > 1 else np.abs(lg[:, 0])
            z_or = [cache.z[j].clone() for j in range(n_hidden)]
            A = cache.conv_out[last].numpy().astype(np.float32)
            sl = slice(b0 - s0, b1 - s0)
            for ei, (e, e_tau) in enumerate(engs):
                for mi, (class_idx, mode) in enumerate(modes):
                    d_cls, d_logits, d_heat, d_z = dev[ei][mi]
                    r = out[ei][mi]
                    over = np.zeros(b1 - b0, bool)
                    for j in range(n_hidden):
                        cache.z[j] = z_or[j]
                        zo = z_or[j].numpy()
                        differ = (zo > 0) != (d_z[j][sl] > 0)
                        r["mask_violations"] += int((differ & (np.abs(zo) > e_tau)).sum())
                        over |= differ.any(axis=1)
                        if differ.any():   # the device's branch of LeakyReLU' for these units (the forward values stay the oracle's)
                            zj = zo.copy()
                            zj[differ] = np.where(d_z[j][sl][differ] > 0, 1e-300, -1e-300)
                            cache.z[j] = torch.as_tensor(zj)
                    ci = o_cls if class_idx is None else np.asarray(class_idx)[b0:b1]
                    cag, _, _ = ocnn.backward(cfg, params, cache, ocnn.top_gradient(cache, ci, mode), through_input=False)
                    o_heat = ogc.gradcam_tail_nhwc(A, cag[last].numpy().astype(np.float32), cfg.input_shape[:2])
                    r["cls_equal"][b0:b1] = d_cls[sl] == o_cls
                    r["logit_err"][b0:b1] = np.abs(d_logits[sl] - lg).max(axis=1)
                    r["heat_err"][b0:b1] = np.abs(d_heat[sl] - o_heat).reshape(b1 - b0, -1).max(axis=1)
                    r["overridden"][b0:b1] = over
                    r["margin"][b0:b1] = margin
                    r["logit_absmax"][b0:b1] = np.abs(lg).max(axis=1)
            for j in range(n_hidden):
                cache.z[j] = z_or[j]
    return out[0] if single else out
