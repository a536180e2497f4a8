from ocpg_b200.functions import MSDeformAttnFunction  # noqa: F401
