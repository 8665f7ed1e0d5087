"""SPVCNN sparse 3D UNet + the FusionTransformer 3D segmentation heads on libft3d.

Topology, channel plan and parameter names follow FusionTransformer/models/spvcnn.py:82-233 (so the
reference's checkpoints load with ``load_state_dict``); the fusion variants follow
models/middle_fusion.py:10-88, early_fusion.py:9-87, late_fusion.py:4-35 and lidar_model.py:4-22.
This module exists so that the hot path can be exercised where /root/reference is absent (GPU box);
the reference's own model files run unmodified after ``install_as_torchsparse()``.
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import cat  # noqa: F401  (re-exported: the reference calls torchsparse.cat)
from .fused import deconv_cat
from . import nn as spnn
from .point_tensor import PointTensor
from .sparse_tensor import SparseTensor
from .voxel_glue import initial_voxelize, point_to_voxel, voxel_to_point

__all__ = ["SPVCNN", "Net3DSeg"]


class BasicConvolutionBlock(nn.Module):
    def __init__(self, inc, outc, ks=3, stride=1, dilation=1):
        super().__init__()
        self.net = nn.Sequential(spnn.Conv3d(inc, outc, kernel_size=ks, dilation=dilation, stride=stride),
                                 spnn.BatchNorm(outc), spnn.ReLU(True))

    def forward(self, x):
        return self.net(x)


class BasicDeconvolutionBlock(nn.Module):
    def __init__(self, inc, outc, ks=3, stride=1):
        super().__init__()
        self.net = nn.Sequential(spnn.Conv3d(inc, outc, kernel_size=ks, stride=stride, transpose=True),
                                 spnn.BatchNorm(outc), spnn.ReLU(True))

    def forward(self, x):
        return self.net(x)


class ResidualBlock(nn.Module):
    def __init__(self, inc, outc, ks=3, stride=1, dilation=1):
        super().__init__()
        self.net = nn.Sequential(
            spnn.Conv3d(inc, outc, kernel_size=ks, dilation=dilation, stride=stride), spnn.BatchNorm(outc),
            spnn.ReLU(True),
            spnn.Conv3d(outc, outc, kernel_size=ks, dilation=dilation, stride=1), spnn.BatchNorm(outc))
        if inc == outc and stride == 1:
            self.downsample = nn.Sequential()
        else:
            self.downsample = nn.Sequential(spnn.Conv3d(inc, outc, kernel_size=1, dilation=1, stride=stride),
                                            spnn.BatchNorm(outc))
        self.relu = spnn.ReLU(True)

    def forward(self, x):
        return self.relu(self.net(x) + self.downsample(x))


def _point_mlp(inc, outc):
    return nn.Sequential(nn.Linear(inc, outc), nn.BatchNorm1d(outc), nn.ReLU(True))


class SPVCNN(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        cr = kwargs.get("cr", 1.0)
        cs = [int(cr * c) for c in (32, 32, 64, 128, 256, 256, 128, 96, 96)]
        self.cs = cs
        if "pres" in kwargs and "vres" in kwargs:
            self.pres, self.vres = kwargs["pres"], kwargs["vres"]
        else:
            self.pres = self.vres = 1

        self.stem = nn.Sequential(
            spnn.Conv3d(4, cs[0], kernel_size=3, stride=1), spnn.BatchNorm(cs[0]), spnn.ReLU(True),
            spnn.Conv3d(cs[0], cs[0], kernel_size=3, stride=1), spnn.BatchNorm(cs[0]), spnn.ReLU(True))
        for i in range(4):   # encoder: stride-2 conv + two residual blocks per stage
            setattr(self, "stage%d" % (i + 1), nn.Sequential(
                BasicConvolutionBlock(cs[i], cs[i], ks=2, stride=2, dilation=1),
                ResidualBlock(cs[i], cs[i + 1], ks=3, stride=1, dilation=1),
                ResidualBlock(cs[i + 1], cs[i + 1], ks=3, stride=1, dilation=1)))
        skips = (cs[3], cs[2], cs[1], cs[0])
        for i in range(4):   # decoder: stride-2 transposed conv, skip concat, two residual blocks
            setattr(self, "up%d" % (i + 1), nn.ModuleList([
                BasicDeconvolutionBlock(cs[4 + i], cs[5 + i], ks=2, stride=2),
                nn.Sequential(ResidualBlock(cs[5 + i] + skips[i], cs[5 + i], ks=3, stride=1, dilation=1),
                              ResidualBlock(cs[5 + i], cs[5 + i], ks=3, stride=1, dilation=1))]))
        self.point_transforms = nn.ModuleList([_point_mlp(cs[0], cs[4]), _point_mlp(cs[4], cs[6]),
                                               _point_mlp(cs[6], cs[8])])
        self.weight_initialization()
        self.dropout = nn.Dropout(0.3, True)
        if os.environ.get("FT3D_FUSE", "1") != "0":
            from .fused import fuse
            fuse(self)                      # conv+BN+ReLU(+shortcut) chains run as fused kernels; modules unchanged

    def weight_initialization(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm1d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def backbone(self, x: SparseTensor, early_feats=None, middle_feats=None, taps=None, plan=None):
        """spvcnn.py:191-233.  ``early_feats`` [N,32] is added at z0 (early_fusion.py:39), ``middle_feats``
        [N,256] at z1 (middle_fusion.py:48); both already passed through their fusion MLP."""
        plan = plan if plan is not None else getattr(x, "plan", None)     # plan.py: geometry built one step ahead
        z = PointTensor(x.F, x.C.float() if plan is None else plan.point_coords)
        x0 = initial_voxelize(z, self.pres, self.vres, plan=plan)
        x0 = self.stem(x0)
        z0 = voxel_to_point(x0, z, nearest=False)
        if early_feats is not None:
            z0.F = z0.F + early_feats

        x1 = self.stage1(point_to_voxel(x0, z0))
        x2 = self.stage2(x1)
        x3 = self.stage3(x2)
        x4 = self.stage4(x3)
        z1 = voxel_to_point(x4, z0)
        z1.F = z1.F + self.point_transforms[0](z0.F)
        if middle_feats is not None:
            z1.F = z1.F + middle_feats

        y1 = point_to_voxel(x4, z1)
        y1.F = self.dropout(y1.F)
        y1 = self.up1[1](deconv_cat(self.up1[0], y1, x3))      # cat([deconv(y), skip]) written in place (fused.py)
        y2 = self.up2[1](deconv_cat(self.up2[0], y1, x2))
        z2 = voxel_to_point(y2, z1)
        z2.F = z2.F + self.point_transforms[1](z1.F)

        y3 = point_to_voxel(y2, z2)
        y3.F = self.dropout(y3.F)
        y3 = self.up3[1](deconv_cat(self.up3[0], y3, x1))
        y4 = self.up4[1](deconv_cat(self.up4[0], y3, x0))
        z3 = voxel_to_point(y4, z2)
        z3.F = z3.F + self.point_transforms[2](z2.F)
        if taps is not None:
            taps.update(x0=x0, x1=x1, x2=x2, x3=x3, x4=x4, y1=y1, y2=y2, y3=y3, y4=y4, z=z, z0=z0, z1=z1,
                        z2=z2, z3=z3)
        return z3.F

    def forward(self, x):
        return self.backbone(x)


class Net3DSeg(SPVCNN):
    """3D branch with segmentation head(s).  ``fusion``: 'none' (lidar_model.py / late_fusion.py),
    'middle' (middle_fusion.py) or 'early' (early_fusion.py)."""

    def __init__(self, num_classes=20, dual_head=False, fusion="none", backbone_3d_kwargs=None):
        super().__init__(**(backbone_3d_kwargs or {}))
        self.fusion = fusion
        if fusion == "middle":
            self.middle_fusion_transform = _point_mlp(96, self.cs[4])
        elif fusion == "early":
            self.early_fusion_transform = _point_mlp(96, 32)
        elif fusion != "none":
            raise ValueError(fusion)
        self.linear = nn.Linear(self.cs[-1], num_classes)
        self.dual_head = dual_head
        if dual_head:
            self.linear2 = nn.Linear(self.cs[-1], num_classes)
        if os.environ.get("FT3D_FUSE", "1") != "0":
            from .fused import fuse
            fuse(self)                      # the fusion MLP added above

    def forward(self, x, img_feats=None, taps=None, plan=None):
        early = middle = None
        if self.fusion == "middle":
            middle = self.middle_fusion_transform(img_feats)
        elif self.fusion == "early":
            early = self.early_fusion_transform(img_feats)
        feats = self.backbone(x, early_feats=early, middle_feats=middle, taps=taps, plan=plan)
        preds = {"lidar_feats": feats, "lidar_seg_logit": self.linear(feats)}
        if self.dual_head:
            preds["lidar_seg_logit2"] = self.linear2(feats)
        return preds
