"""fusiontransformer_b200 -- B200-native (sm_100a) 3D-branch hot path of FusionTransformer.

Public surface = the torchsparse v1.1.0 names the reference imports:

    import fusiontransformer_b200 as torchsparse
    import fusiontransformer_b200.nn as spnn
    import fusiontransformer_b200.nn.functional as spf          # via nn.functional alias
    from fusiontransformer_b200.sparse_tensor import SparseTensor
    from fusiontransformer_b200.point_tensor import PointTensor
    from fusiontransformer_b200.utils import sparse_quantize

``install_as_torchsparse()`` registers those modules under the name ``torchsparse`` so that
FusionTransformer/models/*.py, data/collate.py and the dataloader run unmodified.
All arithmetic runs in libft3d.so (include/ft3d.h); there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import sys

import torch

from .sparse_tensor import SparseTensor
from .point_tensor import PointTensor
from . import functional, nn, utils  # noqa: F401

nn.functional = functional
sys.modules[__name__ + ".nn.functional"] = functional

from .fused import fuse, unfuse  # noqa: E402

__all__ = ["SparseTensor", "PointTensor", "cat", "nn", "utils", "install_as_torchsparse", "fuse", "unfuse"]
__version__ = "0.1.0"


def cat(input_list, dim=1):
    """torchsparse.cat (spvcnn.py:212,216,224,228): concatenate features, keep the first tensor's coords/maps."""
    out = input_list[0]._like(torch.cat([t.F for t in input_list], dim))
    f16 = [getattr(t, "F16", None) for t in input_list]
    if dim == 1 and all(h is not None for h in f16):
        out.F16 = torch.cat(f16, 1)        # keep the tensor-core operand copy alive across the skip concatenation
    return out


def install_as_torchsparse():
    """Make ``import torchsparse...`` resolve to this package (drop-in for the reference's imports)."""
    me = sys.modules[__name__]
    alias = {
        "torchsparse": me,
        "torchsparse.nn": nn,
        "torchsparse.nn.functional": functional,
        "torchsparse.sparse_tensor": sys.modules[__name__ + ".sparse_tensor"],
        "torchsparse.point_tensor": sys.modules[__name__ + ".point_tensor"],
        "torchsparse.utils": utils,
        "torchsparse.utils.kernel_region": sys.modules[__name__ + ".utils.kernel_region"],
        "torchsparse.utils.helpers": sys.modules[__name__ + ".utils.helpers"],
    }
    sys.modules.update(alias)
    return me
