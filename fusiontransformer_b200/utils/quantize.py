"""sparse_quantize -- GPU replacement for torchsparse.utils.sparse_quantize v1.1.0.

Reference call site: FusionTransformer/data/semantic_kitti/semantic_kitti_dataloader.py:231
``inds, _, inverse_map = sparse_quantize(coords, feats, labels, return_index=True, return_invs=True)``.
The voxel key (64-bit multiply-then-xor FNV over the three coordinate columns), the ascending-key order of
``inds``, first-occurrence selection and the inverse map are reproduced bit-exactly by ft3d_quantize
(stable radix sort of (key,row) + head flags + prefix sum) -- see csrc/unique.cu.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


def _to_cuda_i32(coords, quantization_size):
    is_np = isinstance(coords, np.ndarray)
    t = torch.as_tensor(coords)
    if t.dim() != 2 or t.shape[1] != 3:
        raise ValueError("sparse_quantize expects coords of shape [n,3], got %s" % (tuple(t.shape),))
    t = t.cuda(non_blocking=True) if not t.is_cuda else t
    if t.is_floating_point() or quantization_size != 1:
        t = torch.floor(t.double() / quantization_size)
    t = t.to(torch.int32)
    n = t.shape[0]
    c4 = torch.zeros((n, 4), dtype=torch.int32, device=t.device)
    c4[:, :3] = t
    return c4, is_np


def sparse_quantize(coords, feats=None, labels=None, ignore_label=-100, return_index=False,
                    return_invs=False, hash_type="fnv", quantization_size=1):
    if hash_type != "fnv":
        raise ValueError("only hash_type='fnv' is supported (the reference never passes another)")
    use_label = labels is not None
    use_feat = feats is not None
    if not use_label and not use_feat:
        return_index = True
    c4, is_np = _to_cuda_i32(coords, quantization_size)
    inds, invs, _ = ops.quantize(c4, 1)

    def out(t, dtype=torch.int64):
        t = t.to(dtype)
        return t.cpu().numpy() if is_np else t

    inds64 = inds.long()
    if use_label:
        counts = ops.count(invs, inds.numel())
        lab = torch.as_tensor(labels).to(c4.device)[inds64].clone()
        lab[counts > 1] = ignore_label
        lab = lab.cpu().numpy() if is_np else lab
        if return_index:
            return (out(inds), lab, out(invs)) if return_invs else (out(inds), lab)
        disc = c4[inds64, :3]
        f = torch.as_tensor(feats).to(c4.device)[inds64]
        res = (out(disc, torch.int32), f.cpu().numpy() if is_np else f, lab)
        return res + (out(invs),) if return_invs else res
    if return_index:
        return (out(inds), out(invs)) if return_invs else out(inds)
    disc = c4[inds64, :3]
    if use_feat:
        f = torch.as_tensor(feats).to(c4.device)[inds64]
        res = (out(disc, torch.int32), f.cpu().numpy() if is_np else f)
    else:
        res = (out(disc, torch.int32),)
    res = res + (out(invs),) if return_invs else res
    return res if len(res) > 1 else res[0]


def sparse_quantize_batch(points: torch.Tensor, scan_id: torch.Tensor, num_scans: int, scale: float = 20.0,
                          full_scale: int = 4096, rot: torch.Tensor | None = None, transl_u: torch.Tensor | None = None):
    """Device-side a1+a2+a3 for a whole batch: raw points [n,3] f32 in metres with
    their scan ids (ascending, contiguous) -> (coords int32 [m,4] after the bounds filter, kept row ids [m],
    unique first-occurrence rows [U] into the kept set ordered by (scan, key), inverse [m], per-scan counts).
    Follows data/utils/augmentation_3d.py:43-46 and semantic_kitti_dataloader.py:220-231; ``rot`` / ``transl_u`` (the
    per-scan draws of utils/augment.py) switch the augmentation branch (:22-41, :48-51) on."""
    coords, keep = ops.scale_coords(points, scan_id, num_scans, scale, full_scale, rot=rot, transl_u=transl_u)
    kept = torch.nonzero(keep).flatten()
    vc = coords[kept].contiguous()
    inds, invs, scan_counts = ops.quantize(vc, num_scans)
    return vc, kept, inds, invs, scan_counts
