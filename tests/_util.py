"""Shared helpers for the parity tests."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(state):
    h = hashlib.sha256()
    for k in sorted(state):
        h.update(k.encode())
        h.update(state[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def clone_state(state, grad=False, is_trainable=None):
    out = {}
    for k, v in state.items():
        t = v.detach().clone()
        if grad and is_trainable(k) and t.is_floating_point():
            t.requires_grad_(True)
        out[k] = t
    return out


def compare_grads(got: dict, ref: dict, ref_norms: dict = None):
    """Global relative L2 error and the worst per-tensor error measured against the global scale.

    ``ref`` may hold leading slices (64 elements) of large tensors, as the fixtures do; ~10 bias
    tensors per coupling have an analytically zero gradient, so per-tensor errors are
    normalised by max(|ref tensor|, 1e-3 * largest gradient entry overall).
    """
    num = den = 0.0
    gmax = max(float(r.abs().max()) for r in ref.values())
    worst, worst_k = 0.0, None
    for k, r in ref.items():
        g = got[k].detach().cpu()
        gg = g if g.numel() == r.numel() else g.flatten()[: r.numel()]
        gg = gg.reshape(r.shape)
        num += float(((gg - r).double() ** 2).sum())
        den += float((r.double() ** 2).sum())
        e = float((gg - r).abs().max() / max(float(r.abs().max()), 1e-3 * gmax))
        if ref_norms is not None:
            n = ref_norms[k]
            e = max(e, abs(float(g.double().norm()) - n) / max(n, 1e-3 * gmax * max(1.0, g.numel() ** 0.5)))
        if e > worst:
            worst, worst_k = e, k
    return (num / max(den, 1e-300)) ** 0.5, worst, worst_k
