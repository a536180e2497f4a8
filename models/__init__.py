"""Import shim: puts ocpg_b200 at the reference's package path so that the caller's
``from models.ops.modules import MSDeformAttn`` (models/deformable_transformer.py:20) resolves to the
B200 implementation when this repository precedes the reference on ``sys.path``.  Only ``models.ops``
and ``models.deformable_transformer`` (the callers of the path) exist here; the rest of the
reference's ``models`` package is out of scope."""
