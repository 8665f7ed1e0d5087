"""Dispatch of the sparse-convolution arithmetic to libft3d kernels.

Two arithmetic modes, selected by ``FT3D_CONV`` (default ``tc``):
  * ``tc``  -- bf16 operands, fp32 accumulation in TMEM on the tcgen05 tensor cores (csrc/conv_os.cu: forward,
               dgrad, transposed; csrc/conv_pairs_tc.cu: dense k=1 GEMMs and the weight gradient); layers whose
               reduction width is not a multiple of 16 (the 4-channel stem conv) use the fp32 path.
  * ``f32`` -- fp32 CUDA-core kernels (csrc/conv_simt.cu): exact-precision mode for 1e-5 parity checks.
Both consume the same maps; neither has a fallback outside libft3d.
"""
from __future__ import annotations

import os

import torch

from . import ops


_environ = os.environ


def mode() -> str:
    m = _environ.get("FT3D_CONV")
    if m is None:
        return "tc"
    if m not in ("tc", "f32"):
        raise ValueError("FT3D_CONV must be 'tc' or 'f32'")
    return m


WORK_LOG = None   # when a list: one dict per conv launch (kind, pairs, red, ncols, rows) for bench.py's roofline


def _log(kind, kmap, red, ncols, rows):
    if WORK_LOG is not None:
        WORK_LOG.append(dict(kind=kind, pairs=kmap.num_pairs(), red=red, ncols=ncols, rows=rows, K=kmap.K))


def gather_conv(inp, table, kmap, kernel, kflip: bool, w_transposed: bool):
    """Exact fp32 CUDA-core path: out[j] = sum_k inp[table[j,k]] @ (W[k] | W[k]^T); forward, dgrad and transposed
    conv share it (FT3D_CONV=f32, and the layers the tensor-core kernels do not cover)."""
    cin, cout = kernel.shape[-2], kernel.shape[-1]
    red, ncols = (cout, cin) if w_transposed else (cin, cout)
    _log("conv_gather_f32", kmap, red, ncols, table.shape[0])
    w = kernel.detach()
    if w.dim() == 2:
        w = w.unsqueeze(0)
    return ops.conv_gather_f32(inp, table, kmap.K, kflip, w, w_transposed)


def wgrad(feats, gout, kmap, cin: int, cout: int, transpose: bool):
    pairs, offsets = kmap.pairs_padded, kmap.pair_offsets
    max_pairs = kmap.num_pairs()
    ca = 1 if transpose else 0
    _log("conv_wgrad_f32", kmap, cin, cout, max_pairs)
    return ops.conv_wgrad_f32(feats, gout, pairs, offsets, kmap.K, ca, cin, cout, max_pairs)


# ------------------------------------------------------------------------------------------ pair-major tcgen05 path
def pairs_ok(cin: int, cout: int) -> bool:
    """Both the forward (red=cin, ncols=cout), the dgrad (red=cout, ncols=cin) and the wgrad shape must be covered."""
    def gemm_ok(red, ncols):
        return red % 16 == 0 and 16 <= red <= 512 and ncols % 32 == 0 and (32 <= ncols <= 256 or ncols == 384)
    algo = _environ.get("FT3D_CONV_ALGO")
    return (mode() == "tc" and algo in (None, "os", "pairs") and gemm_ok(cin, cout) and gemm_ok(cout, cin)
            and cout <= 256)


def pairs_partial(x16, kmap, kernel, role: str):
    """Pair-major GEMM of one conv: returns (partial rows [L,ncols] f32, the pair-position table that scatters them).
    role: forward | dgrad | transposed | dgrad_transposed (which side of the map is gathered / scattered)."""
    w = kernel.detach()
    K, L = kmap.K, kmap.num_pairs()
    pairs, offsets = kmap.pairs_padded, kmap.pair_offsets
    cin, cout = w.shape[-2], w.shape[-1]
    if role == "forward":
        gcol, ppos, wt, ncols = 0, kmap.ppos, False, cout
    elif role == "dgrad":
        gcol, ppos, wt, ncols = 1, kmap.pposT, True, cin
    elif role == "transposed":       # out = fine rows (pair column 0), gather coarse rows (column 1)
        gcol, ppos, wt, ncols = 1, kmap.pposT, False, cout
    elif role == "dgrad_transposed":
        gcol, ppos, wt, ncols = 0, kmap.ppos, True, cin
    else:
        raise ValueError(role)
    red = cout if wt else cin
    if WORK_LOG is not None:
        WORK_LOG.append(dict(kind="conv_pairs_tc", pairs=L, red=red, ncols=ncols, rows=ppos.shape[0], K=K,
                             rows_in=x16.shape[0]))
    return ops.conv_pairs_tc(x16, pairs, offsets, K, gcol, L, w, wt, owner=kernel), ppos, ncols


def os_enabled() -> bool:
    """FT3D_CONV_ALGO: unset / "os" = output-stationary conv_os for forward, dgrad and transposed convolutions;
    "pairs" = the pair-major GEMM + sorted scatter of csrc/conv_pairs_tc.cu."""
    algo = _environ.get("FT3D_CONV_ALGO")
    return algo is None or algo == "os"


OS_CLUSTER_BY_STRIDE = {1: 1, 2: 1, 4: 1, 8: 1, 16: 1}     # k3 maps: CTAs per cluster sharing the weight blocks


def os_cluster_for_stride(stride: int) -> int:
    m = _environ.get("FT3D_OS_CLUSTER_MAP")          # e.g. "4:2,8:2,16:4"
    if m:
        for item in m.split(","):
            s, c = item.split(":")
            if int(s) == stride:
                return int(c)
    return OS_CLUSTER_BY_STRIDE.get(stride, 1)


def os_tile_rows(kmap, red: int, ncols: int) -> int:
    """Rows of a schedule tile = 128 x the CTAs of the thread-block cluster that shares the tile's weight blocks.
    ``FT3D_OS_CLUSTER`` = 1 | 2 | 4 forces the cluster size; by default it follows the kernel map's tensor stride so
    that every layer on a map uses ONE schedule (the deep, wide layers are the ones bound by re-streaming B_k)."""
    cs = _environ.get("FT3D_OS_CLUSTER")
    if cs is not None and cs != "auto":
        return 128 * int(cs)
    return 128 * getattr(kmap, "os_cluster", 1)


def os_conv(x16, kmap, kernel, role: str, bn=None):
    """Output-stationary tcgen05 convolution of one layer (csrc/conv_os.cu): -> (y f32 [rows, ncols], stat or None).
    ``bn`` = (eps, momentum, running_mean, running_var): BatchNorm training statistics from the epilogue."""
    w = kernel.detach()
    cin, cout = w.shape[-2], w.shape[-1]
    plan, wt, kflip, n_rows = kmap.os_args(role, os_tile_rows(kmap, cin, cout))
    red, ncols = (cout, cin) if wt else (cin, cout)
    if WORK_LOG is not None:
        WORK_LOG.append(dict(kind="conv_os", pairs=kmap.num_pairs(), red=red, ncols=ncols, rows=n_rows, K=kmap.K,
                             rows_in=x16.shape[0], passes=plan.passes()))
    return ops.conv_os(x16, plan, w, wt, kflip, n_rows, owner=kernel, bn=bn)


def pairs_conv(x16, kmap, kernel, role: str):
    partial, ppos, ncols = pairs_partial(x16, kmap, kernel, role)
    return ops.conv_reduce(partial, ppos, ncols)


def pairs_wgrad(x16, g16, kmap, cin: int, cout: int, transpose: bool, into=None, stream=None):
    L = kmap.num_pairs()
    if WORK_LOG is not None:
        WORK_LOG.append(dict(kind="conv_wgrad_pairs_tc", pairs=L, red=cin, ncols=cout, rows=L, K=kmap.K))
    return ops.conv_wgrad_pairs_tc(x16, g16, kmap.pairs_padded, kmap.pair_offsets, kmap.K, 1 if transpose else 0,
                                   cin, cout, L, into=into, stream=stream)


def dense_conv(x16, kernel, w_transposed: bool):
    w = kernel.detach().unsqueeze(0)
    n = x16.shape[0]
    if WORK_LOG is not None:
        WORK_LOG.append(dict(kind="conv_pairs_tc", pairs=n, red=x16.shape[1],
                             ncols=w.shape[1] if w_transposed else w.shape[2], rows=n, K=1, rows_in=n))
    return ops.conv_pairs_tc(x16, None, None, 1, 0, n, w, w_transposed, owner=kernel)[:n]


def dense_wgrad(x16, g16, cin: int, cout: int, into=None):
    n = x16.shape[0]
    if WORK_LOG is not None:
        WORK_LOG.append(dict(kind="conv_wgrad_pairs_tc", pairs=n, red=cin, ncols=cout, rows=n, K=1))
    return ops.conv_wgrad_pairs_tc(x16, g16, None, None, 1, 0, cin, cout, n, into=into).view(cin, cout)
