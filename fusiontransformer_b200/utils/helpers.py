"""torchsparse.utils.helpers v1.1.0 stand-in (wildcard-imported by FusionTransformer/models/utils.py:7 and
spvcnn.py:13).  No ``__all__``: the reference expects ``torch`` and ``np`` to leak through the wildcard."""
import numpy as np
import torch

from .quantize import sparse_quantize  # noqa: F401


def make_tuple(inputs, dimension=3):
    if isinstance(inputs, (list, tuple)):
        assert len(inputs) == dimension
        return tuple(inputs)
    return (inputs,) * dimension
