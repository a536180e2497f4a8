"""Import shim for the reference path ``models.deformable_transformer`` (models/ocpg.py:17 does
``from .deformable_transformer import build_deforamble_transformer``): the classes of that file, re-hosted on the B200
path (ocpg_b200/transformer.py, encoder.py, decoder.py).  Same names, constructor arguments, forward signatures, results."""
from ocpg_b200.decoder import DeformableTransformerDecoder, DeformableTransformerDecoderLayer  # noqa: F401
from ocpg_b200.encoder import DeformableTransformerEncoder, DeformableTransformerEncoderLayer  # noqa: F401
from ocpg_b200.transformer import DeformableTransformer, build_deforamble_transformer  # noqa: F401
