"""ctypes binding of oracle/msda_oracle.c (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Operates on CPU numpy arrays / torch CPU tensors; f32 and f64 only, like the reference op
(ms_deform_attn_cuda.cu:64 dispatches float and double).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile msda_oracle.c into oracle/_ref/libmsda_oracle.so (make; gcc only)."""
    src = os.path.join(_HERE, "msda_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _np(x, dtype):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _dims(value, shapes, loc):
    N, S, M, D = value.shape
    L = shapes.shape[0]
    Lq, P = loc.shape[1], loc.shape[4]
    assert loc.shape == (N, Lq, M, L, P, 2), loc.shape
    return [ctypes.c_int(int(v)) for v in (N, S, M, D, L, Lq, P)]


def forward(value, shapes, start, loc, attn, dtype=np.float64):
    """out[N, Lq, M*D] computed by the C restatement in ``dtype`` (np.float32 | np.float64)."""
    sfx = {np.float32: "f32", np.float64: "f64"}[np.dtype(dtype).type]
    value, loc, attn = _np(value, dtype), _np(loc, dtype), _np(attn, dtype)
    shapes, start = _np(shapes, np.int64), _np(start, np.int64)
    N, S, M, D = value.shape
    out = np.empty((N, loc.shape[1], M * D), dtype=dtype)
    fn = getattr(_load(), f"msda_oracle_forward_{sfx}")
    fn.restype = None
    fn(_ptr(value), _ptr(shapes), _ptr(start), _ptr(loc), _ptr(attn), *_dims(value, shapes, loc), _ptr(out))
    return out


def backward(value, shapes, start, loc, attn, grad_out, dtype=np.float64):
    """(grad_value, grad_loc, grad_attn) by the C restatement."""
    sfx = {np.float32: "f32", np.float64: "f64"}[np.dtype(dtype).type]
    value, loc, attn, grad_out = (_np(t, dtype) for t in (value, loc, attn, grad_out))
    shapes, start = _np(shapes, np.int64), _np(start, np.int64)
    gv, gl, ga = np.empty_like(value), np.empty_like(loc), np.empty_like(attn)
    fn = getattr(_load(), f"msda_oracle_backward_{sfx}")
    fn.restype = None
    fn(_ptr(value), _ptr(shapes), _ptr(start), _ptr(loc), _ptr(attn), _ptr(grad_out),
       *_dims(value, shapes, loc), _ptr(gv), _ptr(gl), _ptr(ga))
    return gv, gl, ga
