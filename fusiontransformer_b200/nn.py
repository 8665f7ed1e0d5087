"""torchsparse.nn v1.1.0 module surface: Conv3d, BatchNorm, ReLU.

Parameter names and shapes match the reference's checkpoints: ``Conv3d.kernel`` is ``[k^3, Cin, Cout]``
(``[Cin, Cout]`` when k == 1); ``BatchNorm`` is an ``nn.BatchNorm1d`` over the rows of ``.F``.
Used by FusionTransformer/models/spvcnn.py:26-31,42-47,57-75,99-102.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import functional as spf
from .sparse_tensor import SparseTensor

__all__ = ["Conv3d", "BatchNorm", "ReLU", "LeakyReLU"]


class Conv3d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilation=1, bias=False, transpose=False):
        super().__init__()
        self.in_channels = self.inc = in_channels
        self.out_channels = self.outc = out_channels
        self.kernel_size = self.ks = kernel_size
        self.k = kernel_size ** 3
        self.stride = self.s = stride
        self.dilation = self.d = dilation
        self.t = transpose
        shape = (self.k, in_channels, out_channels) if self.k > 1 else (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.zeros(*shape))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.init_weight()

    def init_weight(self):
        std = 1.0 / math.sqrt(self.out_channels if self.t else self.in_channels * self.k)
        self.kernel.data.uniform_(-std, std)
        if self.bias is not None:
            self.bias.data.uniform_(-std, std)

    def extra_repr(self):
        return "%d, %d, kernel_size=%d, stride=%d%s" % (self.inc, self.outc, self.ks, self.s,
                                                        ", transpose" if self.t else "")

    def forward(self, inputs: SparseTensor) -> SparseTensor:
        return spf.conv3d(inputs, self.kernel, self.ks, self.bias, self.s, self.d, self.t)


class BatchNorm(nn.BatchNorm1d):
    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__(num_features=num_features, eps=eps, momentum=momentum)

    def forward(self, inputs: SparseTensor) -> SparseTensor:
        return inputs._like(super().forward(inputs.F))


class ReLU(nn.ReLU):
    def forward(self, inputs: SparseTensor) -> SparseTensor:
        return inputs._like(super().forward(inputs.F))


class LeakyReLU(nn.LeakyReLU):
    def forward(self, inputs: SparseTensor) -> SparseTensor:
        return inputs._like(super().forward(inputs.F))
