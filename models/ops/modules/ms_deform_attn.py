"""models.ops.modules.ms_deform_attn -- same import path as the reference file."""
from ocpg_b200.modules.ms_deform_attn import MSDeformAttn, _is_power_of_2  # noqa: F401
