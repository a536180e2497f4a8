"""models.ops.functions.ms_deform_attn_func -- same import path as the reference file."""
from ocpg_b200.functions.ms_deform_attn_func import MSDeformAttnFunction  # noqa: F401
