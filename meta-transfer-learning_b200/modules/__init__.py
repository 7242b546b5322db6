"""API-compatible stand-ins for the reference's ``modules`` package (modules/__init__.py:1-3).

The classes carry the reference's constructor signatures, attribute names, parameter names/shapes and
initialisation draws, so ``state_dict`` keys interchange with reference checkpoints; the arithmetic runs
in libmtl_b200 (see models/asr/transformer.py), not here."""
from .encoder import Encoder  # noqa: F401
from .decoder import Decoder  # noqa: F401
from .discriminator import Discriminator  # noqa: F401
