"""Drop-in for the reference's compiled extension module ``MultiScaleDeformableAttention``.

The reference builds a pybind11 module (models/ops/src/vision.cpp:13-16) exporting

    ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step) -> Tensor
    ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step) -> [grad_value, grad_sampling_loc, grad_attn_weight]

(dispatch: src/ms_deform_attn.h:36-77; CUDA host side: src/cuda/ms_deform_attn_cuda.cu:20-153).  This
module has the same two functions with the same argument meaning and error behaviour, implemented as
a ctypes call into libmsda_sm100.so (include/msda_sm100.h): raw ``data_ptr()``s, the current stream of
the tensors' device, outputs allocated here (the reference allocates them in C++ with at::zeros).

Deliberate supersets of the reference behaviour:
  * ``im2col_step`` is accepted and validated (> 0) but only kept for API parity: the kernels take the
    whole batch in one launch, so the ``batch % min(batch, im2col_step) == 0`` rule (cu:50-52) is gone;
  * bfloat16 ``value`` (with fp32 ``sampling_loc`` / ``attn_weight``) is accepted in addition to
    float32 / float64 (cu:64 dispatches float and double only);
  * CPU tensors raise ``RuntimeError("Not implemented on the CPU")`` exactly like ms_deform_attn.h:54.
"""
from __future__ import annotations

from typing import List

import torch

from . import _lib

_SUFFIX = {torch.float32: "f32", torch.float64: "f64", torch.bfloat16: "bf16"}


def _check_inputs(named, value):
    for name, t in named:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")          # cu:28-32, :93-98
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")                        # ms_deform_attn.h:54
    for name, t in named:
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")                 # cu:34-38, :100-105
        if t.device != value.device:
            raise RuntimeError(f"{name} is on {t.device}, value is on {value.device}")


def _dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
    if value.dim() != 4:
        raise RuntimeError("value must have shape (N, S, M, D)")
    N, S, M, D = value.shape                                                    # cu:40-43
    L = spatial_shapes.shape[0]                                                 # cu:45
    if spatial_shapes.shape != (L, 2) or level_start_index.shape != (L,):
        raise RuntimeError("spatial_shapes must be (L, 2) and level_start_index (L,)")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64 tensors")   # cu:67-68 read int64
    if sampling_loc.dim() != 6:
        raise RuntimeError("sampling_loc must have shape (N, Lq, M, L, P, 2)")
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]                        # cu:47-48
    if sampling_loc.shape != (N, Lq, M, L, P, 2) or attn_weight.shape != (N, Lq, M, L, P):
        raise RuntimeError(f"inconsistent shapes: value {tuple(value.shape)}, sampling_loc "
                           f"{tuple(sampling_loc.shape)}, attn_weight {tuple(attn_weight.shape)}")
    return N, S, M, D, L, Lq, P


def _dtype_suffix(value, sampling_loc, attn_weight):
    sfx = _SUFFIX.get(value.dtype)
    if sfx is None:
        raise RuntimeError(f"ms_deform_attn: unsupported value dtype {value.dtype}")
    want = torch.float32 if sfx == "bf16" else value.dtype
    if sampling_loc.dtype != want or attn_weight.dtype != want:
        raise RuntimeError(f"sampling_loc / attn_weight must be {want} when value is {value.dtype}")
    return sfx


# bf16 value: how grad_value is accumulated.  False (default): fp32 buffer + one conversion pass (error of ONE bf16
# rounding, the tolerance the tests state).  True: packed bf16 reds straight into the bf16 grad_value (half the L2 atomic
# sectors, no buffer, no conversion pass; every red rounds the running sum -- see DESIGN.md for the measured error).
BF16_GRAD_VALUE_DIRECT = False


def set_bf16_grad_value_direct(on: bool) -> None:
    global BF16_GRAD_VALUE_DIRECT
    BF16_GRAD_VALUE_DIRECT = bool(on)


def _stream(device) -> int:
    return _lib.raw_stream(device)                                               # cu:65: current stream


# Optional device timing of the library calls inside a larger program (bench: the operator's share of an
# encoder step).  start_timing() .. stop_timing() brackets every C call with CUDA events on its stream.
_TIMINGS = None


class _Timed:
    def __init__(self, kind, device):
        self.kind, self.device = kind, device

    def __enter__(self):
        if _TIMINGS is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if _TIMINGS is not None and exc[0] is None:
            self.b.record(torch.cuda.current_stream(self.device))
            _TIMINGS.append((self.kind, self.a, self.b))
        return False


def start_timing():
    global _TIMINGS
    _TIMINGS = []


def stop_timing():
    """-> {kind: (calls, total_ms)} for kind in forward / backward (synchronises)."""
    global _TIMINGS
    rec, _TIMINGS = _TIMINGS or [], None
    torch.cuda.synchronize()
    out = {}
    for kind, a, b in rec:
        n, ms = out.get(kind, (0, 0.0))
        out[kind] = (n + 1, ms + a.elapsed_time(b))
    return out


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                           im2col_step: int) -> torch.Tensor:
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)]
    _check_inputs(named, value)
    N, S, M, D, L, Lq, P = _dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    if int(im2col_step) <= 0:
        raise RuntimeError("im2col_step must be positive")
    sfx = _dtype_suffix(value, sampling_loc, attn_weight)
    with _lib.on_device(value.device), _Timed("forward", value.device):
        output = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)   # fully overwritten
        rc = getattr(_lib.lib(), f"msda_forward_{sfx}")(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
            attn_weight.data_ptr(), N, S, M, D, L, Lq, P, output.data_ptr(), _stream(value.device))
    _lib.check(rc, "ms_deform_attn_forward")
    return output


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step: int) -> List[torch.Tensor]:
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), ("grad_output", grad_output)]
    _check_inputs(named, value)
    N, S, M, D, L, Lq, P = _dims(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    if int(im2col_step) <= 0:
        raise RuntimeError("im2col_step must be positive")
    if grad_output.dtype != value.dtype or grad_output.numel() != N * Lq * M * D:
        raise RuntimeError("grad_output must have value's dtype and N*Lq*M*D elements")
    sfx = _dtype_suffix(value, sampling_loc, attn_weight)
    with _lib.on_device(value.device), _Timed("backward", value.device):
        grad_loc = torch.empty_like(sampling_loc)      # every element is written by the kernel
        grad_attn = torch.empty_like(attn_weight)
        st = _stream(value.device)
        L_ = _lib.lib()
        if sfx == "bf16":
            direct = BF16_GRAD_VALUE_DIRECT and L <= 4
            acc = None if direct else torch.empty(value.shape, dtype=torch.float32, device=value.device)   # zero-filled by the call
            grad_value = torch.empty_like(value)
            rc = L_.msda_backward_bf16(
                grad_output.data_ptr(), value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                sampling_loc.data_ptr(), attn_weight.data_ptr(), N, S, M, D, L, Lq, P,
                None if direct else acc.data_ptr(), grad_value.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(), st)
        else:
            grad_value = torch.empty_like(value)                                       # zero-filled by the call
            rc = getattr(L_, f"msda_backward_{sfx}")(
                grad_output.data_ptr(), value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                sampling_loc.data_ptr(), attn_weight.data_ptr(), N, S, M, D, L, Lq, P,
                grad_value.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(), st)
    _lib.check(rc, "ms_deform_attn_backward")
    return [grad_value, grad_loc, grad_attn]


# ------------------------------------------------------------------------------------------------
# Fused module path (not in the reference extension): softmax of the logits and the
# reference-point / offset arithmetic of MSDeformAttn.forward (ms_deform_attn.py:100-110) run inside
# the kernels.  Same checks and error behaviour as above.
# ------------------------------------------------------------------------------------------------
def fused_supported(value, n_levels: int, n_points: int) -> bool:
    """True if the fused kernels cover this layout (the tiled sm_100a kernels: 32 channels per head ...)."""
    if value.dtype not in (torch.float32, torch.bfloat16) or value.dim() != 4:
        return False
    return _lib.lib().msda_kernel_plan(value.element_size(), value.shape[2], value.shape[3], n_levels, n_points) == 1


def _fused_dims(value, spatial_shapes, level_start_index, offsets, logits, ref):
    if offsets.dim() != 6 or logits.dim() != 4 or ref.dim() != 4:
        raise RuntimeError("offsets must be (N, Lq, M, L, P, 2), logits (N, Lq, M, L*P), reference_points (N, Lq, L, 2|4)")
    N, Lq, M_, L_, P, _ = offsets.shape
    attn_like = logits.view(N, Lq, M_, L_, P) if logits.shape == (N, Lq, M_, L_ * P) else logits
    dims = _dims(value, spatial_shapes, level_start_index, offsets, attn_like)
    ref_dim = ref.shape[-1]
    if ref.shape != (N, Lq, L_, ref_dim) or ref_dim not in (2, 4):
        raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(ref_dim))  # :112-113
    for name, t in (("offsets", offsets), ("logits", logits), ("reference_points", ref)):
        if t.dtype != torch.float32:
            raise RuntimeError(f"{name} must be float32")
    return dims + (ref_dim,)


def ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index, offsets, logits, reference_points,
                                 im2col_step: int, emit: bool = True):
    """-> (output, sampling_locations | None, attention_weights | None)."""
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("offsets", offsets), ("logits", logits), ("reference_points", reference_points)]
    _check_inputs(named, value)
    N, S, M, D, L, Lq, P, ref_dim = _fused_dims(value, spatial_shapes, level_start_index, offsets, logits, reference_points)
    if int(im2col_step) <= 0:
        raise RuntimeError("im2col_step must be positive")
    sfx = {torch.float32: "f32", torch.bfloat16: "bf16"}.get(value.dtype)
    if sfx is None:
        raise RuntimeError(f"fused path: unsupported value dtype {value.dtype}")
    with _lib.on_device(value.device), _Timed("forward", value.device):
        output = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        loc = torch.empty_like(offsets) if emit else None
        attn = torch.empty((N, Lq, M, L, P), dtype=torch.float32, device=value.device) if emit else None
        rc = getattr(_lib.lib(), f"msda_fused_forward_{sfx}")(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), offsets.data_ptr(),
            logits.data_ptr(), reference_points.data_ptr(), ref_dim, N, S, M, D, L, Lq, P, output.data_ptr(),
            loc.data_ptr() if emit else None, attn.data_ptr() if emit else None, _stream(value.device))
    _lib.check(rc, "ms_deform_attn_fused_forward")
    return output, loc, attn


def ms_deform_attn_fused_backward(value, spatial_shapes, level_start_index, offsets, logits, reference_points,
                                  grad_output, im2col_step: int, need_grad_loc: bool = False):
    """-> [grad_value, grad_offsets, grad_logits, grad_sampling_locations | None]."""
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("offsets", offsets), ("logits", logits), ("reference_points", reference_points),
             ("grad_output", grad_output)]
    _check_inputs(named, value)
    N, S, M, D, L, Lq, P, ref_dim = _fused_dims(value, spatial_shapes, level_start_index, offsets, logits, reference_points)
    if grad_output.dtype != value.dtype or grad_output.numel() != N * Lq * M * D:
        raise RuntimeError("grad_output must have value's dtype and N*Lq*M*D elements")
    sfx = {torch.float32: "f32", torch.bfloat16: "bf16"}.get(value.dtype)
    if sfx is None:
        raise RuntimeError(f"fused path: unsupported value dtype {value.dtype}")
    with _lib.on_device(value.device), _Timed("backward", value.device):
        g_off = torch.empty_like(offsets)
        g_logits = torch.empty_like(logits)
        g_loc = torch.empty_like(offsets) if need_grad_loc else None
        g_loc_ptr = g_loc.data_ptr() if need_grad_loc else None
        st = _stream(value.device)
        L_ = _lib.lib()
        common = (grad_output.data_ptr(), value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                  offsets.data_ptr(), logits.data_ptr(), reference_points.data_ptr(), ref_dim, N, S, M, D, L, Lq, P)
        grad_value = torch.empty_like(value)
        if sfx == "bf16":
            direct = BF16_GRAD_VALUE_DIRECT and L <= 4
            acc = None if direct else torch.empty(value.shape, dtype=torch.float32, device=value.device)
            rc = L_.msda_fused_backward_bf16(*common, None if direct else acc.data_ptr(), grad_value.data_ptr(), g_off.data_ptr(),
                                             g_logits.data_ptr(), g_loc_ptr, st)
        else:
            rc = L_.msda_fused_backward_f32(*common, grad_value.data_ptr(), g_off.data_ptr(), g_logits.data_ptr(),
                                            g_loc_ptr, st)
    _lib.check(rc, "ms_deform_attn_fused_backward")
    return [grad_value, g_off, g_logits, g_loc]
