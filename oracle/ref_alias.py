"""`torchsparse` module tree backed by the ORACLE operators, so that the reference's own model files execute
unmodified on the CPU.  TEST INFRASTRUCTURE ONLY.

The reference imports torchsparse v1.1.0 at FusionTransformer/models/spvcnn.py:7-13, models/utils.py:1-7 and
models/middle_fusion.py:2-4 (``import torchsparse``, ``torchsparse.nn as spnn``, ``torchsparse.nn.functional as
spf``, ``torchsparse.sparse_tensor.SparseTensor``, ``torchsparse.point_tensor.PointTensor``, and wildcard imports
of ``torchsparse.utils.kernel_region`` / ``torchsparse.utils.helpers`` through which models/utils.py receives the
name ``torch``).  ``install()`` registers modules of those names whose symbols are the oracle's restatements
(oracle/ts_ops.py, oracle/ft_glue.py); ``load_reference_models()`` then imports the reference's
``models/spvcnn.py`` (SPVCNN, ResidualBlock, ...), ``models/utils.py`` (initial_voxelize, point_to_voxel,
voxel_to_point) and ``models/middle_fusion.py`` (Net3DSeg.backbone_forward_pass, :32-88) from /root/reference.

What this pins: the TOPOLOGY and GLUE the oracle restates in ft_glue.py (layer order, channel plan, skip
concatenations, point<->voxel call order, caching keys, parameter names) against the reference's own code, executed,
not read.  What it cannot pin: the operator arithmetic itself (torchsparse is absent), which both sides take from
oracle/ts_ops.py.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

from . import ft_glue as og
from . import ts_ops as ts

REF_ROOT = "/root/reference"
_NAMES = ["torchsparse", "torchsparse.nn", "torchsparse.nn.functional", "torchsparse.sparse_tensor",
          "torchsparse.point_tensor", "torchsparse.utils", "torchsparse.utils.kernel_region",
          "torchsparse.utils.helpers"]


def install():
    """Register the oracle-backed ``torchsparse`` modules in sys.modules (replacing any other alias)."""
    mods = {n: types.ModuleType(n) for n in _NAMES}
    top, nn_, fn = mods["torchsparse"], mods["torchsparse.nn"], mods["torchsparse.nn.functional"]
    top.SparseTensor, top.PointTensor, top.cat = ts.SparseTensor, ts.PointTensor, ts.cat
    top.nn, top.utils = nn_, mods["torchsparse.utils"]
    top.sparse_tensor, top.point_tensor = mods["torchsparse.sparse_tensor"], mods["torchsparse.point_tensor"]
    nn_.Conv3d, nn_.BatchNorm, nn_.ReLU, nn_.functional = og.Conv3d, og.BatchNorm, og.ReLU, fn
    for name in ("sphash", "sphashquery", "spcount", "spvoxelize", "spdevoxelize", "calc_ti_weights", "conv3d"):
        setattr(fn, name, getattr(ts, name))
    mods["torchsparse.sparse_tensor"].SparseTensor = ts.SparseTensor
    mods["torchsparse.point_tensor"].PointTensor = ts.PointTensor
    kr, hp, ut = (mods["torchsparse.utils.kernel_region"], mods["torchsparse.utils.helpers"],
                  mods["torchsparse.utils"])
    kr.KernelRegion, kr.torch = ts.KernelRegion, torch         # models/utils.py gets `torch` through the wildcard
    hp.torch = torch
    ut.kernel_region, ut.helpers, ut.sparse_quantize = kr, hp, ts.sparse_quantize
    for n, m in mods.items():
        sys.modules[n] = m
    return mods


def uninstall():
    for n in _NAMES:
        sys.modules.pop(n, None)
    for n in [n for n in sys.modules if n == "FusionTransformer" or n.startswith("FusionTransformer.")]:
        del sys.modules[n]


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "FusionTransformer", "models", "spvcnn.py"))


def load_reference_models():
    """-> (spvcnn module, utils module, middle_fusion module) of the REFERENCE, imported from /root/reference on
    the oracle-backed alias.  middle_fusion.py:5 imports the 2D branch (timm, absent here, out of scope): that one
    module is replaced by a stub holding the two class names it imports."""
    if not available():
        raise RuntimeError("the reference tree is not present on this box")
    uninstall()
    install()
    stub = types.ModuleType("FusionTransformer.models.image_models")
    stub.Net2DSeg = stub.Net2DBillinear = type("_Unused2DBranch", (), {})
    sys.path.insert(0, REF_ROOT)
    try:
        pkg = types.ModuleType("FusionTransformer")                      # skip the package __init__ side effects
        pkg.__path__ = [os.path.join(REF_ROOT, "FusionTransformer")]
        sys.modules["FusionTransformer"] = pkg
        mpkg = types.ModuleType("FusionTransformer.models")
        mpkg.__path__ = [os.path.join(REF_ROOT, "FusionTransformer", "models")]
        sys.modules["FusionTransformer.models"] = mpkg
        sys.modules["FusionTransformer.models.image_models"] = stub
        spv = importlib.import_module("FusionTransformer.models.spvcnn")
        utl = importlib.import_module("FusionTransformer.models.utils")
        mid = importlib.import_module("FusionTransformer.models.middle_fusion")
    finally:
        sys.path.pop(0)
    return spv, utl, mid
