"""ocpg_b200 -- B200-native (sm_100a) multi-scale deformable attention for TJUMMG/OCPG.

A drop-in for the reference's ``models/ops`` package and, around it, for the callers of that path in
``models/deformable_transformer.py`` (section 8f):

    from ocpg_b200 import MSDeformAttn, MSDeformAttnFunction          # same API as models.ops.{modules,functions}
    import ocpg_b200.MultiScaleDeformableAttention as MSDA            # same functions as the pybind11 module
    from ocpg_b200.transformer import DeformableTransformer           # encoder.py / decoder.py / flatten.py / epilogue.py

Python/PyTorch host code calls hand-written CUDA kernels through the C ABI of include/msda_sm100.h
(ctypes); no Triton, no multi-backend dispatch, no CPU fallback.
"""
from ._strict import set_strict, is_strict, fallback_counts  # noqa: F401
from ._lib import build, lib, launch_count, set_option, source_fingerprint, LIB_PATH  # noqa: F401
from .functions import MSDeformAttnFunction, MSDeformAttnFusedFunction  # noqa: F401
from .modules import MSDeformAttn  # noqa: F401
from . import MultiScaleDeformableAttention  # noqa: F401

__all__ = ["MSDeformAttn", "MSDeformAttnFunction", "MSDeformAttnFusedFunction", "MultiScaleDeformableAttention", "build", "lib"]
