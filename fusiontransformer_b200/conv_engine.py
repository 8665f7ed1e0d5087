"""Dispatch of the sparse-convolution arithmetic to libft3d kernels.

Two arithmetic modes, selected by ``FT3D_CONV`` (default ``tc``):
  * ``tc``  -- bf16 operands, fp32 accumulation in TMEM on the tcgen05 tensor cores (csrc/conv_tc.cu);
               layers whose reduction width is not a multiple of 16 (the 4-channel stem conv) use the fp32 path.
  * ``f32`` -- fp32 CUDA-core kernels (csrc/conv_simt.cu): exact-precision mode for 1e-5 parity checks.
Both consume the same maps; neither has a fallback outside libft3d.
"""
from __future__ import annotations

import os

import torch

from . import ops


def mode() -> str:
    m = os.environ.get("FT3D_CONV", "tc")
    if m not in ("tc", "f32"):
        raise ValueError("FT3D_CONV must be 'tc' or 'f32'")
    return m


def _tc_ok(red: int, ncols: int) -> bool:
    return (mode() == "tc" and red % 16 == 0 and 16 <= red <= 512 and ncols % 32 == 0
            and (32 <= ncols <= 256 or ncols == 384))


WORK_LOG = None   # when a list: one dict per conv launch (kind, pairs, red, ncols, rows) for bench.py's roofline


def _log(kind, kmap, red, ncols, rows):
    if WORK_LOG is not None:
        WORK_LOG.append(dict(kind=kind, pairs=kmap.num_pairs(), red=red, ncols=ncols, rows=rows, K=kmap.K))


def gather_conv(inp, table, kmap, kernel, kflip: bool, w_transposed: bool):
    """out[j] = sum_k inp[table[j,k]] @ (W[k] | W[k]^T); forward, dgrad and transposed conv share it."""
    cin, cout = kernel.shape[-2], kernel.shape[-1]
    red, ncols = (cout, cin) if w_transposed else (cin, cout)
    _log("conv_gather_tc" if _tc_ok(red, ncols) else "conv_gather_f32", kmap, red, ncols, table.shape[0])
    w = kernel.detach()
    if w.dim() == 2:
        w = w.unsqueeze(0)
    if _tc_ok(red, ncols):
        return ops.conv_gather_tc(inp, table, kmap.K, kflip, w, w_transposed, owner=kernel)
    return ops.conv_gather_f32(inp, table, kmap.K, kflip, w, w_transposed)


def wgrad(feats, gout, kmap, cin: int, cout: int, transpose: bool):
    pairs, offsets = kmap.pairs_padded, kmap.pair_offsets
    max_pairs = kmap.num_pairs()
    ca = 1 if transpose else 0
    tc = mode() == "tc" and cin % 16 == 0 and 16 <= cin <= 512 and cout % 32 == 0 and 32 <= cout <= 256
    _log("conv_wgrad_tc" if tc else "conv_wgrad_f32", kmap, cin, cout, max_pairs)
    if tc:
        return ops.conv_wgrad_tc(feats, gout, pairs, offsets, kmap.K, ca, cin, cout, max_pairs)
    return ops.conv_wgrad_f32(feats, gout, pairs, offsets, kmap.K, ca, cin, cout, max_pairs)
