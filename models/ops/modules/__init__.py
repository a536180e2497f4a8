from ocpg_b200.modules import MSDeformAttn  # noqa: F401
