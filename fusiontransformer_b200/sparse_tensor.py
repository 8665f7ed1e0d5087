"""SparseTensor -- container mirroring torchsparse.SparseTensor v1.1.0.

Constructed by the reference at FusionTransformer/data/collate.py:67 (``SparseTensor(coords=, feats=)``)
and FusionTransformer/models/utils.py:29,59.  Besides the reference-visible attributes
(``F, C, s, coord_maps, kernel_maps``) it carries ``tables``: the per-stride coordinate hash tables
that libft3d reuses across every map build of a forward pass (shared by reference, like coord_maps).
"""
from __future__ import annotations

import torch

__all__ = ["SparseTensor"]


class SparseTensor:
    def __init__(self, feats, coords, stride=1):
        self._F16 = None
        self.F = feats
        self.C = coords
        self.s = stride
        self.coord_maps = {}
        self.kernel_maps = {}
        self.tables = {}

    # ``F16``: bf16 copy of ``F`` written by the fused conv+BN+ReLU epilogue (csrc/bn.cu) -- the operand the next
    # tensor-core convolution gathers.  Assigning ``F`` drops it, so it can never go stale.
    @property
    def F(self):
        return self._F

    @F.setter
    def F(self, value):
        self._F = value
        self._F16 = None

    @property
    def F16(self):
        return self._F16

    @F16.setter
    def F16(self, value):
        self._F16 = value

    def check(self):
        if self.s not in self.coord_maps:
            self.coord_maps[self.s] = self.C

    def _like(self, feats):
        t = SparseTensor(feats, self.C, self.s)
        t.coord_maps, t.kernel_maps, t.tables = self.coord_maps, self.kernel_maps, self.tables
        return t

    def cuda(self):
        assert isinstance(self.F, torch.Tensor) and isinstance(self.C, torch.Tensor)
        self.F = self.F.cuda(non_blocking=True)
        self.C = self.C.cuda(non_blocking=True)
        return self

    def detach(self):
        assert isinstance(self.F, torch.Tensor) and isinstance(self.C, torch.Tensor)
        self.F = self.F.detach()
        self.C = self.C.detach()
        return self

    def to(self, device, non_blocking=True):
        assert isinstance(self.F, torch.Tensor) and isinstance(self.C, torch.Tensor)
        self.F = self.F.to(device, non_blocking=non_blocking)
        self.C = self.C.to(device, non_blocking=non_blocking)
        return self

    def __add__(self, other):
        return self._like(self.F + other.F)
