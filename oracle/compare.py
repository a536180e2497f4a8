"""Comparison helpers for the parity tests (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Error metric: max|a-b| / max|b| per tensor, against the fp64 oracle fed the
same fp32-valued inputs.

``grad_sampling_loc`` is discontinuous where a sample's pixel coordinate ``loc*W - 0.5`` crosses an
integer (bilinear interpolation is piecewise), so an fp32 implementation and the fp64 oracle may put
a point that lies within rounding distance of a cell boundary into different cells; both are right.
``boundary_mask`` marks those points so the test can exclude them (and report how many).
"""
from __future__ import annotations

import numpy as np


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().float().cpu().numpy() if str(x.dtype) == "torch.bfloat16" else x.detach().cpu().numpy()
    return np.asarray(x)


def rel_err(a, b) -> float:
    """max|a-b| / max|b| (0/0 -> 0)."""
    a, b = _np(a).astype(np.float64), _np(b).astype(np.float64)
    den = np.abs(b).max() if b.size else 0.0
    num = np.abs(a - b).max() if b.size else 0.0
    return float(num / den) if den > 0 else float(num)


def boundary_mask(loc, shapes, eps: float = 1e-4):
    """Boolean (N,Lq,M,L,P) array: True where the point's x or y pixel coordinate (computed in fp64
    from the given locations) is within ``eps`` px of an integer, i.e. of a bilinear cell boundary."""
    loc = _np(loc).astype(np.float64)
    shapes = _np(shapes).astype(np.float64)
    L = shapes.shape[0]
    W = shapes[:, 1].reshape(1, 1, 1, L, 1)
    H = shapes[:, 0].reshape(1, 1, 1, L, 1)
    x = loc[..., 0] * W - 0.5
    y = loc[..., 1] * H - 0.5
    near = lambda t: np.abs(t - np.round(t)) < eps
    return near(x) | near(y)
