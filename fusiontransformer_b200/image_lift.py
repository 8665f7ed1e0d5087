"""2D -> 3D lift straight from the low-resolution image-feature map.

The reference head (FusionTransformer/models/image_models_billinear.py:8-23) runs ``Conv2d 1x1 -> ReLU ->
BatchNorm2d`` on the ``[B, C, h, w]`` token map (24 x 24) and then ``nn.Upsample(size=(H, W))`` -- default mode
'nearest' -- to ``[B, 96, 370, 1226]`` (174 MB per sample), only for ``get_img_feats`` (:111-124) to read one pixel per
LiDAR point.  ``lift_nearest`` reads that pixel from the low-resolution map with PyTorch's legacy nearest rule
``src = min(floor(dst * float(in) / float(out)), in - 1)`` (fp32 arithmetic): forward values are
bit-identical to ``lift(upsample(x))`` and the upsampled map is never written; the gradient lands in the small map.
``BilinearLiftHead`` is ``BilinearModule`` with the same parameters / state_dict plus that fused path.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import functional

__all__ = ["nearest_source_index", "lift_nearest", "BilinearLiftHead"]


def nearest_source_index(dst: torch.Tensor, in_size: int, out_size: int) -> torch.Tensor:
    """Source index ATen's upsample_nearest uses for destination index ``dst`` (legacy 'nearest'): the scale is the
    fp32 quotient ``float(in) / float(out)``, the product is taken in fp32, floored, and clamped to ``in - 1``."""
    scale = torch.tensor(float(in_size), dtype=torch.float32, device=dst.device) / \
        torch.tensor(float(out_size), dtype=torch.float32, device=dst.device)
    src = torch.floor(dst.to(torch.float32) * scale).to(torch.int64)
    return src.clamp_(max=in_size - 1)


def lift_nearest(src_map: torch.Tensor, img_indices, out_size, batch_index: torch.Tensor | None = None) -> torch.Tensor:
    """``lift(nn.Upsample(out_size)(src_map), img_indices)`` without the upsampled map.

    src_map [B, C, h, w] f32 (CUDA); img_indices: the reference's list of B ``[N_i, 2]`` (row, col) arrays in
    ``out_size = (H, W)`` pixel coordinates, or one ``[N, 2]`` tensor with ``batch_index`` [N]."""
    H, W = int(out_size[0]), int(out_size[1])
    h, w = src_map.shape[-2], src_map.shape[-1]
    dev = src_map.device
    if isinstance(img_indices, (list, tuple)):
        sizes = [len(a) for a in img_indices]
        rc = torch.from_numpy(np.concatenate([np.asarray(a).reshape(-1, 2) for a in img_indices]).astype(np.int64)).to(dev)
        batch_index = torch.repeat_interleave(torch.arange(len(sizes), device=dev), torch.tensor(sizes, device=dev))
    else:
        rc = img_indices.to(dev)
        if batch_index is None:
            raise ValueError("lift_nearest: a single index tensor needs batch_index")
    if rc.numel() and (int(rc[:, 0].max()) >= H or int(rc[:, 1].max()) >= W or int(rc.min()) < 0):
        raise IndexError("lift_nearest: img_indices outside the (H, W) = (%d, %d) image" % (H, W))
    src_rc = torch.stack([nearest_source_index(rc[:, 0], h, H), nearest_source_index(rc[:, 1], w, W)], 1)
    return functional.lift(src_map, src_rc.to(torch.int32), batch_index.to(torch.int32))


class BilinearLiftHead(nn.Module):
    """``BilinearModule`` (image_models_billinear.py:8-23): same submodules and state_dict keys (``stem.0`` Conv2d 1x1,
    ``stem.2`` BatchNorm2d); ``forward`` is the reference's (materialises the upsampled map), ``lift`` is the fused
    path the 3D branch needs."""

    def __init__(self, in_features, out_features, interpolation_output_size):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(in_features, out_features, kernel_size=1), nn.ReLU(True),
                                  nn.BatchNorm2d(out_features))
        self.up = nn.Upsample(interpolation_output_size)
        self.size = tuple(interpolation_output_size)

    def forward(self, x):
        return self.up(self.stem(x))

    def lift(self, x, img_indices, batch_index=None):
        return lift_nearest(self.stem(x), img_indices, self.size, batch_index)
