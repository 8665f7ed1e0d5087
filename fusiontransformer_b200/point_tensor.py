"""PointTensor -- container mirroring torchsparse.PointTensor v1.1.0 (FusionTransformer/models/spvcnn.py:193,
models/utils.py:88-100): point features/coordinates plus the per-stride point<->voxel map caches."""
from __future__ import annotations

import torch

__all__ = ["PointTensor"]


class PointTensor:
    def __init__(self, feat, coords, idx_query=None, weights=None):
        self.F = feat
        self.C = coords
        self.idx_query = idx_query if idx_query is not None else {}
        self.weights = weights if weights is not None else {}
        self.additional_features = {"idx_query": {}, "counts": {}}

    def cuda(self):
        self.F = self.F.cuda(non_blocking=True)
        self.C = self.C.cuda(non_blocking=True)
        return self

    def detach(self):
        self.F = self.F.detach()
        self.C = self.C.detach()
        return self

    def to(self, device, non_blocking=True):
        self.F = self.F.to(device, non_blocking=non_blocking)
        self.C = self.C.to(device, non_blocking=non_blocking)
        return self

    def __add__(self, other):
        t = PointTensor(self.F + other.F, self.C, self.idx_query, self.weights)
        t.additional_features = self.additional_features
        return t
