"""Oracle restatement of the reference's own hot-path glue and SPVCNN topology.

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see oracle/__init__.py).

Follows (file:line in /root/reference/FusionTransformer):
  models/utils.py:15-35   initial_voxelize      models/utils.py:40-63  point_to_voxel
  models/utils.py:68-106  voxel_to_point        models/spvcnn.py:22-233 SPVCNN
  models/middle_fusion.py:10-88, models/early_fusion.py:9-87, models/lidar_model.py:4-22
  models/image_models_billinear.py:111-124 (2D->3D lift)
  data/semantic_kitti/semantic_kitti_dataloader.py:216-238, data/utils/augmentation_3d.py:43-46
  data/collate.py:36-67, data/utils/validate.py:10-11
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ts_ops as ts
from .ts_ops import PointTensor, SparseTensor


# ----------------------------------------------------------------- dataloader side
def voxelize_scan(points: np.ndarray, scale: int = 20, full_scale: int = 4096):
    """augmentation_3d.py:43-46 (no augmentation) + dataloader :220-231.

    Returns (coords_all int64 [m,3] after the bounds filter, keep mask [n],
    unique_inds [u], inverse_map [m]).
    """
    coords = points * scale
    coords = coords - coords.min(0)
    coords = coords.astype(np.int64)
    keep = (coords.min(1) >= 0) * (coords.max(1) < full_scale)
    vc = coords[keep]
    inds, _, inv = ts.sparse_quantize(vc, np.zeros((len(vc), 1), np.float32),
                                      np.zeros(len(vc), np.int64),
                                      return_index=True, return_invs=True)
    return vc, keep, inds, inv


def augment_draws(noisy_rot=0.0, flip_x=0.0, flip_y=0.0, rot_z=0.0, transl=False, rng=np.random):
    """The random numbers of one call of augmentation_3d.py:22-51, drawn in the reference's order from ``rng`` (the
    global numpy generator by default, as the reference uses it): -> (rot float32 [3,3] | None, u float64 [3] | None)
    with ``u`` the uniform factors of the random translation (they are drawn AFTER the rotation's numbers; the offset
    itself depends on the rotated points, see ``augment_and_scale``)."""
    rot = None
    if noisy_rot > 0 or flip_x > 0 or flip_y > 0 or rot_z > 0:
        rot = np.eye(3, dtype=np.float32)
        if noisy_rot > 0:                       # :25-26  noise on every element (float64 draws added into the f32 matrix)
            rot += rng.randn(3, 3) * noisy_rot
        if flip_x > 0:                          # :28-29  sign of the x axis
            rot[0][0] *= rng.randint(0, 2) * 2 - 1
        if flip_y > 0:                          # :31-32
            rot[1][1] *= rng.randint(0, 2) * 2 - 1
        if rot_z > 0:                           # :34-39  rotation about the up axis, composed on the right
            theta = rng.rand() * rot_z
            rz = np.array([[np.cos(theta), -np.sin(theta), 0], [np.sin(theta), np.cos(theta), 0], [0, 0, 1]],
                          dtype=np.float32)
            rot = rot.dot(rz)
    u = rng.rand(3) if transl else None         # :50
    return rot, u


def augment_and_scale(points: np.ndarray, scale, full_scale, rot=None, u=None):
    """augmentation_3d.py:40-51 given the draws of ``augment_draws``: rotate (float32 matrix product), scale, move to
    the positive octant, optionally translate inside the receptive field.  Returns float32 coordinates [n,3]."""
    if rot is not None:
        points = points.dot(rot)                # :40
    coords = points * scale                     # :43
    coords -= coords.min(0)                     # :46
    if u is not None:                           # :48-51 (float64 offset added into the float32 array)
        offset = np.clip(full_scale - coords.max(0) - 0.001, a_min=0, a_max=None) * u
        coords += offset
    return coords


def collate(scans):
    """collate.py:36-67: append the batch index column and concatenate.

    ``scans``: list of dicts with 'coords' [u,3] int64 and 'feats' [u,4] f32.
    """
    locs, feats = [], []
    for b, s in enumerate(scans):
        c = torch.from_numpy(np.asarray(s["coords"]))
        locs.append(torch.cat([c, torch.full((c.shape[0], 1), b, dtype=torch.int64)], 1))
        feats.append(torch.from_numpy(np.asarray(s["feats"])))
    return SparseTensor(coords=torch.cat(locs, 0), feats=torch.cat(feats, 0))


def map_sparse_to_org(x, inverse_map):
    """validate.py:10-11."""
    return x[inverse_map]


def lift(feature_map: torch.Tensor, img_indices) -> torch.Tensor:
    """image_models_billinear.py:117-124: per-sample NHWC gather at (row, col), concatenated."""
    out = []
    for b in range(feature_map.shape[0]):
        idx = torch.as_tensor(np.asarray(img_indices[b])).long()
        out.append(feature_map.permute(0, 2, 3, 1)[b][idx[:, 0], idx[:, 1]])
    return torch.cat(out, 0)


# ----------------------------------------------------------------- models/utils.py
def initial_voxelize(z: PointTensor, init_res, after_res) -> SparseTensor:
    new_float_coord = torch.cat([(z.C[:, :3] * init_res) / after_res, z.C[:, -1].view(-1, 1)], 1)
    pc_hash = ts.sphash(torch.floor(new_float_coord).int())
    sparse_hash = torch.unique(pc_hash)
    idx_query = ts.sphashquery(pc_hash, sparse_hash)
    counts = ts.spcount(idx_query.int(), len(sparse_hash))
    inserted_coords = ts.spvoxelize(torch.floor(new_float_coord), idx_query, counts)
    inserted_coords = torch.round(inserted_coords).int()
    inserted_feat = ts.spvoxelize(z.F, idx_query, counts)
    new_tensor = SparseTensor(inserted_feat, inserted_coords, 1)
    new_tensor.check()
    z.additional_features["idx_query"][1] = idx_query
    z.additional_features["counts"][1] = counts
    z.C = new_float_coord
    return new_tensor


def _strided_point_coords(z: PointTensor, s: int) -> torch.Tensor:
    return torch.cat([torch.floor(z.C[:, :3] / s).int() * s, z.C[:, -1].int().view(-1, 1)], 1)


def point_to_voxel(x: SparseTensor, z: PointTensor) -> SparseTensor:
    cache = z.additional_features
    if cache["idx_query"].get(x.s) is None:
        pc_hash = ts.sphash(_strided_point_coords(z, x.s))
        idx_query = ts.sphashquery(pc_hash, ts.sphash(x.C))
        counts = ts.spcount(idx_query.int(), x.C.shape[0])
        cache["idx_query"][x.s], cache["counts"][x.s] = idx_query, counts
    else:
        idx_query, counts = cache["idx_query"][x.s], cache["counts"][x.s]
    new_tensor = SparseTensor(ts.spvoxelize(z.F, idx_query, counts), x.C, x.s)
    new_tensor.coord_maps, new_tensor.kernel_maps = x.coord_maps, x.kernel_maps
    return new_tensor


def voxel_to_point(x: SparseTensor, z: PointTensor, nearest=False) -> PointTensor:
    if z.idx_query.get(x.s) is None or z.weights.get(x.s) is None:
        off = ts.KernelRegion(2, x.s, 1).get_kernel_offset()
        old_hash = ts.sphash(_strided_point_coords(z, x.s), off)
        idx_query = ts.sphashquery(old_hash, ts.sphash(x.C))
        weights = ts.calc_ti_weights(z.C, idx_query, scale=x.s).transpose(0, 1).contiguous()
        idx_query = idx_query.transpose(0, 1).contiguous()
        if nearest:
            weights[:, 1:] = 0.0
            idx_query[:, 1:] = -1
        z.idx_query[x.s], z.weights[x.s] = idx_query, weights
    new_feat = ts.spdevoxelize(x.F, z.idx_query[x.s], z.weights[x.s])
    new_tensor = PointTensor(new_feat, z.C, idx_query=z.idx_query, weights=z.weights)
    new_tensor.additional_features = z.additional_features
    return new_tensor


# ----------------------------------------------------------------- spnn modules
class Conv3d(nn.Module):
    def __init__(self, inc, outc, kernel_size=3, stride=1, dilation=1, bias=False, transpose=False):
        super().__init__()
        self.ks, self.k, self.s, self.d, self.t = kernel_size, kernel_size ** 3, stride, dilation, transpose
        shape = (self.k, inc, outc) if self.k > 1 else (inc, outc)
        self.kernel = nn.Parameter(torch.zeros(*shape))
        ts.conv_weight_init(self.kernel.data, inc, outc, self.k, transpose)
        self.bias = None

    def forward(self, x):
        return ts.conv3d(x, self.kernel, self.ks, self.bias, self.s, self.d, self.t)


class BatchNorm(nn.BatchNorm1d):
    def forward(self, x):
        t = SparseTensor(super().forward(x.F), x.C, x.s)
        t.coord_maps, t.kernel_maps = x.coord_maps, x.kernel_maps
        return t


class ReLU(nn.ReLU):
    def forward(self, x):
        t = SparseTensor(torch.relu(x.F), x.C, x.s)
        t.coord_maps, t.kernel_maps = x.coord_maps, x.kernel_maps
        return t


class Linear(nn.Linear):
    """nn.Linear whose GEMM follows ts.OPERAND_DTYPE (same parameters / state_dict as nn.Linear)."""

    def forward(self, x):
        return ts.linear(x, self.weight, self.bias)


def _conv_bn_relu(inc, outc, ks, stride, transpose=False):
    return nn.Sequential(Conv3d(inc, outc, ks, stride=stride, transpose=transpose), BatchNorm(outc), ReLU(True))


class _Block(nn.Module):
    """BasicConvolutionBlock / BasicDeconvolutionBlock (spvcnn.py:22-50): attribute ``net``."""

    def __init__(self, inc, outc, ks, stride, transpose=False):
        super().__init__()
        self.net = _conv_bn_relu(inc, outc, ks, stride, transpose)

    def forward(self, x):
        return self.net(x)


class ResidualBlock(nn.Module):
    """spvcnn.py:53-79."""

    def __init__(self, inc, outc, ks=3, stride=1):
        super().__init__()
        self.net = nn.Sequential(Conv3d(inc, outc, ks, stride=stride), BatchNorm(outc), ReLU(True),
                                 Conv3d(outc, outc, ks, stride=1), BatchNorm(outc))
        self.downsample = nn.Sequential() if (inc == outc and stride == 1) else \
            nn.Sequential(Conv3d(inc, outc, 1, stride=stride), BatchNorm(outc))
        self.relu = ReLU(True)

    def forward(self, x):
        return self.relu(self.net(x) + self.downsample(x))


class SPVCNN(nn.Module):
    """spvcnn.py:82-233 with identical parameter names (state_dict interchangeable)."""

    def __init__(self, **kwargs):
        super().__init__()
        cr = kwargs.get("cr", 1.0)
        cs = [int(cr * x) for x in [32, 32, 64, 128, 256, 256, 128, 96, 96]]
        self.cs = cs
        self.pres = kwargs.get("pres", 1) if ("pres" in kwargs and "vres" in kwargs) else 1
        self.vres = kwargs.get("vres", 1) if ("pres" in kwargs and "vres" in kwargs) else self.pres
        self.stem = nn.Sequential(Conv3d(4, cs[0], 3), BatchNorm(cs[0]), ReLU(True),
                                  Conv3d(cs[0], cs[0], 3), BatchNorm(cs[0]), ReLU(True))
        for i in range(4):
            setattr(self, "stage%d" % (i + 1), nn.Sequential(
                _Block(cs[i], cs[i], 2, 2), ResidualBlock(cs[i], cs[i + 1]), ResidualBlock(cs[i + 1], cs[i + 1])))
        skips = [cs[3], cs[2], cs[1], cs[0]]
        for i in range(4):
            setattr(self, "up%d" % (i + 1), nn.ModuleList([
                _Block(cs[4 + i], cs[5 + i], 2, 2, transpose=True),
                nn.Sequential(ResidualBlock(cs[5 + i] + skips[i], cs[5 + i]), ResidualBlock(cs[5 + i], cs[5 + i]))]))
        self.point_transforms = nn.ModuleList([
            nn.Sequential(Linear(cs[0], cs[4]), nn.BatchNorm1d(cs[4]), nn.ReLU(True)),
            nn.Sequential(Linear(cs[4], cs[6]), nn.BatchNorm1d(cs[6]), nn.ReLU(True)),
            nn.Sequential(Linear(cs[6], cs[8]), nn.BatchNorm1d(cs[8]), nn.ReLU(True))])
        for m in self.modules():
            if isinstance(m, nn.BatchNorm1d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self.dropout = nn.Dropout(0.3, True)

    def backbone(self, x: SparseTensor, early=None, middle=None, taps=None):
        """spvcnn.py:191-233; ``early``/``middle`` are the fusion adds of early_fusion.py:39 /
        middle_fusion.py:48 (already passed through their MLPs by the caller)."""
        z = PointTensor(x.F, x.C.float())
        x0 = initial_voxelize(z, self.pres, self.vres)
        x0 = self.stem(x0)
        z0 = voxel_to_point(x0, z, nearest=False)
        if early is not None:
            z0.F = z0.F + early
        x1 = point_to_voxel(x0, z0)
        x1 = self.stage1(x1)
        x2 = self.stage2(x1)
        x3 = self.stage3(x2)
        x4 = self.stage4(x3)
        z1 = voxel_to_point(x4, z0)
        z1.F = z1.F + self.point_transforms[0](z0.F)
        if middle is not None:
            z1.F = z1.F + middle
        y1 = point_to_voxel(x4, z1)
        y1.F = self.dropout(y1.F)
        y1 = self.up1[0](y1)
        y1 = ts.cat([y1, x3])
        y1 = self.up1[1](y1)
        y2 = self.up2[0](y1)
        y2 = ts.cat([y2, x2])
        y2 = self.up2[1](y2)
        z2 = voxel_to_point(y2, z1)
        z2.F = z2.F + self.point_transforms[1](z1.F)
        y3 = point_to_voxel(y2, z2)
        y3.F = self.dropout(y3.F)
        y3 = self.up3[0](y3)
        y3 = ts.cat([y3, x1])
        y3 = self.up3[1](y3)
        y4 = self.up4[0](y3)
        y4 = ts.cat([y4, x0])
        y4 = self.up4[1](y4)
        z3 = voxel_to_point(y4, z2)
        z3.F = z3.F + self.point_transforms[2](z2.F)
        if taps is not None:
            taps.update(x0=x0, x1=x1, x2=x2, x3=x3, x4=x4, y1=y1, y2=y2, y3=y3, y4=y4,
                        z0=z0, z1=z1, z2=z2, z3=z3, z=z)
        return z3.F

    def forward(self, x):
        return self.backbone(x)


class Net3DSeg(SPVCNN):
    """3D branch + heads.  fusion in {'none','late','middle','early'}
    (lidar_model.py:4-22, late_fusion.py:4-35, middle_fusion.py:10-88, early_fusion.py:9-87)."""

    def __init__(self, num_classes=20, dual_head=False, fusion="middle", backbone_3d_kwargs=None):
        super().__init__(**(backbone_3d_kwargs or {}))
        self.fusion = fusion
        if fusion == "middle":
            self.middle_fusion_transform = nn.Sequential(Linear(96, self.cs[4]), nn.BatchNorm1d(self.cs[4]), nn.ReLU(True))
        elif fusion == "early":
            self.early_fusion_transform = nn.Sequential(Linear(96, 32), nn.BatchNorm1d(32), nn.ReLU(True))
        self.linear = nn.Linear(self.cs[-1], num_classes)
        self.dual_head = dual_head
        if dual_head:
            self.linear2 = nn.Linear(self.cs[-1], num_classes)

    def forward(self, x, img_feats=None, taps=None):
        early = middle = None
        if self.fusion == "middle":
            middle = self.middle_fusion_transform(img_feats)
        elif self.fusion == "early":
            early = self.early_fusion_transform(img_feats)
        feats = self.backbone(x, early=early, middle=middle, taps=taps)
        preds = {"lidar_feats": feats, "lidar_seg_logit": self.linear(feats)}
        if self.dual_head:
            preds["lidar_seg_logit2"] = self.linear2(feats)
        return preds
