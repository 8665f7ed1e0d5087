"""Fused point<->voxel glue: drop-in equivalents of FusionTransformer/models/utils.py.

``initial_voxelize`` (:15-35), ``point_to_voxel`` (:40-63) and ``voxel_to_point`` (:68-106) keep the
reference's signatures, caches (``z.additional_features``, ``z.idx_query``, ``z.weights``) and results, but
each map build is ONE libft3d launch against the per-stride coordinate table instead of the reference's
hash -> table rebuild -> query -> count -> ~30 elementwise kernels.  The reference's own utils.py also runs
unmodified on the operator-level API in ``functional`` (tests/test_reference_topology.py executes it on the oracle
alias; tests/test_gpu_ops.py holds these fused versions to the oracle's results).
"""
from __future__ import annotations

import torch

from . import ops
from .functional import _Devoxelize, _Voxelize, _table_for
from .ops import CoordTable
from .point_tensor import PointTensor
from .sparse_tensor import SparseTensor

__all__ = ["initial_voxelize", "point_to_voxel", "voxel_to_point"]


def initial_voxelize(z: PointTensor, init_res, after_res, plan=None) -> SparseTensor:
    if plan is not None:
        return _initial_voxelize_planned(z, init_res, after_res, plan)
    new_float_coord = torch.cat([(z.C[:, :3] * init_res) / after_res, z.C[:, -1].view(-1, 1)], 1)
    coords = torch.floor(new_float_coord).int()
    pc_hash = ops.hash_coords(coords)
    # sorted unique == torch.unique(pc_hash); inverse == sphashquery(pc_hash, sparse_hash); counts == spcount
    sparse_hash, idx_query, counts, first = ops.unique_sorted(pc_hash)
    # every point of a voxel shares the same floor()ed coordinate, so mean+round == the first member's row
    inserted_coords = ops.gather_rows_i32(coords, first)
    inserted_feat = _Voxelize.apply(z.F.float().contiguous(), idx_query, counts)
    new_tensor = SparseTensor(inserted_feat, inserted_coords, 1)
    new_tensor.check()
    new_tensor.tables[1] = CoordTable(sparse_hash)
    z.additional_features["idx_query"][1] = idx_query
    z.additional_features["counts"][1] = counts
    z.C = new_float_coord
    return new_tensor


def _initial_voxelize_planned(z: PointTensor, init_res, after_res, plan) -> SparseTensor:
    """Same result as above with every integer structure taken from a GeometryPlan (plan.py) built ahead of time."""
    if init_res != after_res:
        raise ValueError("a GeometryPlan is built for pres == vres (the reference's configuration)")
    new_tensor = SparseTensor(_Voxelize.apply(z.F.float().contiguous(), plan.idx_query, plan.counts),
                              plan.coord_maps[1], 1)
    new_tensor.coord_maps = dict(plan.coord_maps)
    new_tensor.kernel_maps = dict(plan.kernel_maps)
    new_tensor.tables = dict(plan.tables)
    z.additional_features["idx_query"][1] = plan.idx_query
    z.additional_features["counts"][1] = plan.counts
    for s, (idx, cnt) in plan.p2v.items():
        z.additional_features["idx_query"][s] = idx
        z.additional_features["counts"][s] = cnt
    for s, (idx, w) in plan.v2p.items():
        z.idx_query[s] = idx
        z.weights[s] = w
    z.C = plan.point_coords
    return new_tensor


def point_to_voxel(x: SparseTensor, z: PointTensor) -> SparseTensor:
    cache = z.additional_features
    if cache is None or cache.get("idx_query") is None or cache["idx_query"].get(x.s) is None:
        table = _table_for(x, x.s, x.C)
        idx_query, counts = ops.p2v_build(z.C.contiguous(), x.s, table, x.C.shape[0])
        cache["idx_query"][x.s] = idx_query
        cache["counts"][x.s] = counts
    else:
        idx_query, counts = cache["idx_query"][x.s], cache["counts"][x.s]
    idx32 = idx_query if idx_query.dtype == torch.int32 else idx_query.int()
    return x._like(_Voxelize.apply(z.F.contiguous(), idx32, counts))


def voxel_to_point(x: SparseTensor, z: PointTensor, nearest: bool = False) -> PointTensor:
    if z.idx_query is None or z.weights is None or z.idx_query.get(x.s) is None or z.weights.get(x.s) is None:
        table = _table_for(x, x.s, x.C)
        idx_query, weights = ops.v2p_build(z.C.contiguous(), x.s, table)
        if nearest:
            weights[:, 1:] = 0.0
            idx_query[:, 1:] = -1
        z.idx_query[x.s] = idx_query
        z.weights[x.s] = weights
    idx_query, weights = z.idx_query[x.s], z.weights[x.s]
    idx32 = idx_query if idx_query.dtype == torch.int32 else idx_query.int()
    new_feat = _Devoxelize.apply(x.F.contiguous(), idx32.contiguous(), weights)
    new_tensor = PointTensor(new_feat, z.C, idx_query=z.idx_query, weights=z.weights)
    new_tensor.additional_features = z.additional_features
    return new_tensor
