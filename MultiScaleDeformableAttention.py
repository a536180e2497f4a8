"""Top-level ``MultiScaleDeformableAttention`` module, the name the reference's
functions/ms_deform_attn_func.py:18 imports (there: a pybind11 extension built by models/ops/setup.py).
Re-exports the ctypes-backed implementation."""
from ocpg_b200.MultiScaleDeformableAttention import ms_deform_attn_forward, ms_deform_attn_backward  # noqa: F401
